"""The handful of `warp` calls sim/physics/phystwin.py makes around the simulator
(:28-30 init / ScopedTimer / set_module_options, :517 capture_launch,
:525,530 to_torch), mapped onto real2sim_eval_b200.physics.SpringMassSystemWarp.
Not a Warp re-implementation."""
import torch


def init():
    return None


def set_module_options(options):
    return None


class ScopedTimer:
    enabled = False

    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def to_torch(a):
    return a if isinstance(a, torch.Tensor) else a.torch()


_graph_owner = {}


def capture_launch(graph):
    """phystwin.py:515-517 replays `simulator.graph`; our `graph` attribute is a
    token owned by a SpringMassSystemWarp whose step() is already one launch."""
    owner = getattr(graph, "owner", None)
    if owner is None:
        raise RuntimeError("capture_launch: unknown graph token")
    owner.step()
