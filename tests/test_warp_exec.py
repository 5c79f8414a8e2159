"""CPU: the warp-lang interpreter the physics goldens are made with (oracle/warp_exec.py), checked on toy kernels
that use each construct of the reference's spring_mass_warp.py: float32 rounding of scalar / vec3 arithmetic,
`i, j = wp.tid()` on a 1-D launch, torch tensors passed for array parameters, atomics, capture + replay with the
launch-time scalar frozen into the node, memory shared with torch, and the restated built-ins' contracts."""
import numpy as np
import pytest
import torch

from oracle import warp_exec as wp

F32 = np.float32


@wp.kernel(enable_backward=False)
def _axpy(x: wp.array(dtype=wp.vec3), y: wp.array(dtype=wp.vec3), a: float, out: wp.array(dtype=wp.vec3)):
    tid = wp.tid()
    out[tid] = x[tid] * a + y[tid] * 0.1


@wp.kernel(enable_backward=False)
def _copy2d(data: wp.array2d(dtype=wp.vec3), origin: wp.array2d(dtype=wp.vec3)):
    i, j = wp.tid()
    origin[i][j] = data[i][j]


@wp.kernel(enable_backward=False)
def _scatter(idx: wp.array(dtype=wp.vec2i), f: wp.array(dtype=wp.vec3), step: int):
    tid = wp.tid()
    wp.atomic_add(f, idx[tid][0], wp.vec3(1.0, 0.0, float(step)))
    wp.atomic_sub(f, idx[tid][1], wp.vec3(1.0, 0.0, 0.0))


def test_arithmetic_rounds_to_float32_like_the_generated_code():
    x = wp.array(np.array([[0.1, 0.2, 0.3]]), dtype=wp.vec3)
    y = wp.array(np.array([[1.0, 2.0, 3.0]]), dtype=wp.vec3)
    out = wp.zeros(1, dtype=wp.vec3)
    wp.launch(_axpy, dim=1, inputs=[x, y, 1.0 / 3.0], outputs=[out])
    a = F32(1.0 / 3.0)
    want = np.array([F32(F32(F32(v) * a) + F32(F32(w) * F32(0.1))) for v, w in zip((0.1, 0.2, 0.3), (1.0, 2.0, 3.0))])
    assert out.numpy().dtype == np.float32 and np.array_equal(out.numpy()[0], want)
    assert wp.length(wp.vec3(3.0, 4.0, 0.0)) == F32(5.0) and isinstance(wp.exp(-1.5e-4), np.float32)
    assert wp.clamp(F32(2.5), low=0.0, high=2.0) == F32(2.0) and wp.max(F32(1e-9), 1e-6) == F32(1e-6)
    assert np.array_equal(wp.normalize(wp.vec3(0.0, 0.0, 0.0)), np.zeros(3, F32))


def test_two_index_tid_on_a_one_dimensional_launch_and_torch_arguments():
    """set_mesh_interactive launches copy_2dvec3 with dim = len(centre table) and passes torch tensors
    (SMW:785-790): j is 0 and a (S, 3) tensor is an (S, 1) array of vec3."""
    src = torch.arange(12, dtype=torch.float32).reshape(4, 3)
    dst = wp.zeros((4, 1), dtype=wp.vec3)
    wp.launch(_copy2d, dim=len(src), inputs=[src], outputs=[dst])
    assert np.array_equal(dst.numpy()[:, 0], src.numpy())
    src2 = torch.arange(24, dtype=torch.float32).reshape(4, 2, 3)
    dst2 = wp.zeros((4, 2), dtype=wp.vec3)
    wp.launch(_copy2d, dim=(4, 2), inputs=[src2], outputs=[dst2])
    assert np.array_equal(dst2.numpy(), src2.numpy())


def test_capture_records_and_replay_executes_with_frozen_scalars():
    idx = wp.from_torch(torch.tensor([[0, 1], [1, 2]], dtype=torch.int32), dtype=wp.vec2i)
    f = wp.zeros(3, dtype=wp.vec3)
    with wp.ScopedCapture() as cap:
        for i in range(2):
            f.zero_()
            wp.launch(_scatter, dim=2, inputs=[idx, f, i])
    assert not f.numpy().any(), "a stream capture does not execute"
    wp.capture_launch(cap.graph)
    assert np.array_equal(f.numpy(), np.array([[1, 0, 1], [0, 0, 1], [-1, 0, 0]], np.float32))   # step 1's node ran last
    wp.capture_launch(cap.graph)
    assert np.array_equal(f.numpy()[:, 0], [1, 0, -1]), "every replay starts from the recorded zero_()"


def test_from_torch_shares_memory_and_to_torch_views_it():
    t = torch.zeros(5, 3)
    a = wp.from_torch(t, dtype=wp.vec3)
    a[2] = wp.vec3(1.0, 2.0, 3.0)
    assert t[2].tolist() == [1.0, 2.0, 3.0] and a.shape == (5,)
    v = a[2]
    v[0] = 9.0
    assert t[2, 0] == 1.0, "reading an element gives a value (register copy), not a view"
    assert wp.to_torch(a).data_ptr() == t.data_ptr()


def test_hash_grid_contract():
    """Cell = C-truncation of p / cell_width, +2^20, mod dim; build sorts by cell then id; a query yields every point
    of every cell overlapping the box, x fastest."""
    pts = wp.array(np.array([[0.03, 0.0, 0.0], [-0.01, 0.0, 0.0], [0.01, 0.0, 0.0], [0.30, 0.0, 0.0]], np.float32), dtype=wp.vec3)
    g = wp.HashGrid(128, 128, 128)
    g.build(pts, 0.025)
    order = [wp.hash_grid_point_id(g.id, k) for k in range(4)]
    assert order == [1, 2, 0, 3], "-0.01 and 0.01 share the double-wide cell 0 (truncation toward zero), in id order"
    near = sorted(wp.hash_grid_query(g.id, wp.vec3(0.0, 0.0, 0.0), 0.025))
    assert near == [0, 1, 2] and 3 not in near


def test_mesh_query_contract():
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    f = np.array([[0, 2, 1], [0, 1, 3], [0, 3, 2], [1, 2, 3]], np.int32)      # outward-facing tetrahedron
    m = wp.Mesh(points=wp.array(v, dtype=wp.vec3), indices=wp.array(f.reshape(-1), dtype=int))
    q = wp.mesh_query_point_sign_winding_number(m.id, wp.vec3(0.2, 0.2, -0.05), max_dist=0.02, accuracy=3.0, threshold=0.6)
    assert not q.result, "farther than max_dist: no hit"
    q = wp.mesh_query_point_sign_winding_number(m.id, wp.vec3(0.2, 0.2, -0.01), max_dist=0.02, accuracy=3.0, threshold=0.6)
    p = wp.mesh_eval_position(m.id, q.face, q.u, q.v)
    assert q.result and q.face == 0 and q.sign == F32(1.0) and np.allclose(p, [0.2, 0.2, 0.0], atol=1e-7)
    q = wp.mesh_query_point_sign_winding_number(m.id, wp.vec3(0.2, 0.2, 0.01), max_dist=0.02, accuracy=3.0, threshold=0.6)
    assert q.result and q.face == 0 and q.sign == F32(-1.0), "inside: winding number ~1 > threshold"


def test_restated_mesh_query_agrees_with_independent_geometry():
    """warp-lang's native library is not available, so `mesh_query_point_sign_winding_number` is restated (parity
    unpinned).  Independent geometry narrows what is unpinned to Warp's tie rule and its far-field approximation: on
    the closed finger mesh the restated query returns the minimum over the triangles of a closest distance computed
    another way (plane projection if it falls inside the triangle, else the nearest of the three edge segments --
    not Ericson's region walk), and its inside / outside sign equals ray-casting parity."""
    from real2sim_eval_b200 import synth
    g = synth.make_gripper(center=(0.0, 0.0, 0.0), gap=0.03)
    v, f = g.verts[:24].astype(np.float64), g.faces[:44]            # the left finger: a closed 44-triangle box
    m = wp.Mesh(points=wp.array(g.verts[:24], dtype=wp.vec3), indices=wp.array(f.reshape(-1).astype(np.int32), dtype=int))
    rng = np.random.default_rng(21)
    lo, hi = v.min(0) - 0.01, v.max(0) + 0.01
    pts = rng.uniform(lo, hi, (300, 3))

    def seg(p, a, b):
        t = np.clip(np.dot(p - a, b - a) / np.dot(b - a, b - a), 0.0, 1.0)
        return np.linalg.norm(p - (a + t * (b - a)))

    def tri_dist(p, a, b, c):
        n = np.cross(b - a, c - a)
        n /= np.linalg.norm(n)
        q = p - np.dot(p - a, n) * n
        inside = all(np.dot(np.cross(e1 - e0, q - e0), n) >= 0 for e0, e1 in ((a, b), (b, c), (c, a)))
        return abs(np.dot(p - a, n)) if inside else min(seg(p, a, b), seg(p, b, c), seg(p, c, a))

    def inside_by_parity(p):
        d = np.array([0.3713, 0.5821, 0.7233])                       # generic direction: no edge or vertex hits
        hits = 0
        for a, b, c in v[f]:
            e1, e2 = b - a, c - a
            h = np.cross(d, e2)
            det = np.dot(e1, h)
            if abs(det) < 1e-15:
                continue
            s = p - a
            u = np.dot(s, h) / det
            q = np.cross(s, e1)
            w = np.dot(d, q) / det
            hits += int(u >= 0 and w >= 0 and u + w <= 1 and np.dot(e2, q) / det > 0)
        return hits % 2 == 1

    n_in = 0
    for p in pts:
        q = wp.mesh_query_point_sign_winding_number(m.id, wp.vec3(*p.astype(np.float32)), max_dist=1.0, accuracy=3.0,
                                                    threshold=0.6)
        assert q.result
        hit = np.asarray(wp.mesh_eval_position(m.id, q.face, q.u, q.v), np.float64)
        want = min(tri_dist(p, *v[t]) for t in f)
        assert abs(np.linalg.norm(p.astype(np.float32) - hit) - want) < 2e-6
        inside = inside_by_parity(p)
        assert (q.sign < 0) == inside
        n_in += inside
    assert 20 < n_in < 280, "points on both sides of the surface"
