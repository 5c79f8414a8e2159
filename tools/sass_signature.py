"""Floating-point instruction signature of a kernel in libr2s.so (no GPU needed): the multiset of FP opcodes with their
modifiers and operand shapes (registers anonymised), as `cuobjdump -sass` prints them.  The bit-identity of
preprocess_kernel with the reference build depends on WHICH multiplies ptxas fuses with WHICH adds, and that follows
the surrounding code; the signature changes when the pairing does, so a CPU-only test can say "re-run the GPU
bit-identity tests" before a GPU is involved.
    python tools/sass_signature.py [kernel-substring] [--lib path] [--write tests/golden/sass_signature.json]"""
import argparse
import hashlib
import json
import os
import re
import shutil
import subprocess
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP = ("FFMA", "FMUL", "FADD", "MUFU", "DFMA", "DMUL", "DADD", "F2F", "FMNMX", "FCHK", "FRND", "F2I", "I2F")


def signature(lib, kernel):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([exe, "-sass", lib], capture_output=True, text=True, check=True).stdout
    on, ops = False, Counter()
    for line in out.splitlines():
        if "Function :" in line:
            on = kernel in line
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?)\s*;", line)
        if not (on and m):
            continue
        t = m.group(1).split()
        t = t[1:] if t[0].startswith("@") else t
        if t[0].split(".")[0] in FP:
            shape = re.sub(r"\bR\d+\b", "R", " ".join(t)).replace(".reuse", "")
            shape = re.sub(r"\bP\d\b", "P", shape)
            ops[shape] += 1
    by_op = Counter()
    for k, v in ops.items():
        by_op[k.split()[0].split(".")[0]] += v
    digest = hashlib.sha256(json.dumps(sorted(ops.items())).encode()).hexdigest()[:16]
    return {"kernel": kernel, "fp_instructions": sum(ops.values()), "by_opcode": dict(sorted(by_op.items())), "digest": digest}


def toolchain():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    out = subprocess.run([exe, "--version"], capture_output=True, text=True).stdout
    m = re.search(r"V(\d+\.\d+\.\d+)", out)
    return m.group(1) if m else "unknown"


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("kernel", nargs="?", default="preprocess_kernel")
    ap.add_argument("--lib", default=os.path.join(ROOT, "real2sim_eval_b200", "libr2s.so"))
    ap.add_argument("--write")
    a = ap.parse_args()
    sig = dict(signature(a.lib, a.kernel), nvcc=toolchain())
    print(json.dumps(sig, indent=1))
    if a.write:
        json.dump(sig, open(a.write, "w"), indent=1)
