"""Extract the T-block PhysTwin graph shipped with the reference into a small fixture.

Run in the build container (needs /root/reference):
    python tests/golden/make_tblock_fixture.py
Writes tests/golden/tblock.npz {x, v, springs, rest, spring_Y}.  The source is
/root/reference/experiments/utils/T_final_state.pkl (pickled CUDA tensors; the only
data artefact the reference ships, consumed by calculate_success_T.py:51-53).
"""
import io
import os
import pickle

import numpy as np
import torch

SRC = "/root/reference/experiments/utils/T_final_state.pkl"


class _CpuUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "torch.storage" and name == "_load_from_bytes":
            return lambda b: torch.load(io.BytesIO(b), map_location="cpu", weights_only=False)
        return super().find_class(module, name)


def main():
    with open(SRC, "rb") as f:
        d = _CpuUnpickler(f).load()
    out = dict(
        x=d["renderer"]["x"].numpy().astype(np.float32),
        v=d["renderer"]["v"].numpy().astype(np.float32),
        springs=d["model"]["init_springs"].numpy().astype(np.int32),
        rest=d["model"]["init_rest_lengths"].numpy().astype(np.float32),
        spring_Y=d["model"]["init_spring_Y"].numpy().astype(np.float32),
    )
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tblock.npz")
    np.savez_compressed(dst, **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()}, os.path.getsize(dst))


if __name__ == "__main__":
    main()
