"""CPU oracles (TEST INFRASTRUCTURE).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; the
product path (real2sim_eval_b200/) never does."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(ref: bool = True, quiet: bool = True) -> None:
    """Compile the C restatements (and oracle/_ref when /root/reference is mounted)."""
    target = ["all"] if ref else ["oracles"]
    subprocess.run(["make", "-C", _HERE] + target, check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def lib_path(name: str) -> str:
    p = os.path.join(_HERE, "_build", name)
    if not os.path.exists(p):
        build(ref=False)
    return p


def ref_lib_path() -> str | None:
    p = os.path.join(_HERE, "_ref", "libref_raster.so")
    return p if os.path.exists(p) else None
