#!/usr/bin/env python3
"""Mandatory build step (run by __graft_entry__.build() and `make -C gpu-benchmarking_b200/csrc check`, and again by
tests/test_cabi_cpu.py): scan the SASS of the shipped libb200fe.so for the one structural property the
programmatic-dependent-launch path rests on.

The basis matrices live in a per-device __constant__ bank that a tiny fill kernel rewrites before every per-call
operator; the operator is launched as the fill's programmatic dependent and may be resident while the fill runs.
That is only safe if NO instruction of the operator can read the bank (SASS operand `c[0x3][...]`) before
`griddepcontrol.wait` (SASS `ACQBULK`) has returned.  ptxas treats __constant__ data as immutable and will hoist
uniform loads above the wait inside one function, so every such kernel must (1) wait within its first few
instructions and (2) reach its first bank read only through a real CALL to a non-inlined body.  A kernel that
violates this would multiply by the previous call's basis -- silently.  This script fails the build instead.

Exit status 0 = every waiting kernel has the wait -> call -> bank-read order; 1 = violations (listed).
"""
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_SO = os.path.join(ROOT, "gpu-benchmarking_b200", "libb200fe.so")


def scan(so=DEFAULT_SO):
    """-> (number of kernels that wait, list of (function, state) violating the order)"""
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        raise FileNotFoundError("cuobjdump")
    proc = subprocess.Popen([exe, "-sass", so], stdout=subprocess.PIPE, text=True)
    fn, idx, state = None, 0, {}
    fn_re, ins_re = re.compile(r"\s+Function : (\S+)"), re.compile(r"\s+/\*[0-9a-f]{4}\*/")
    for line in proc.stdout:
        m = fn_re.match(line)
        if m:
            fn, idx = m.group(1), 0
            state[fn] = {"acq": None, "call": None, "bank": None}
            continue
        if fn and ins_re.match(line):
            idx += 1
            st = state[fn]
            if st["acq"] is None and "ACQBULK" in line:
                st["acq"] = idx
            if st["call"] is None and "CALL" in line:
                st["call"] = idx
            if st["bank"] is None and "c[0x3]" in line:
                st["bank"] = idx
    if proc.wait() != 0:
        raise RuntimeError("cuobjdump failed")
    waits, bad = 0, []
    for fn, st in state.items():
        if st["acq"] is None:
            continue
        waits += 1
        if st["acq"] > 8 or st["call"] is None or (st["bank"] is not None and st["bank"] < st["call"]):
            bad.append((fn, st))
    return waits, bad


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else DEFAULT_SO
    waits, bad = scan(so)
    print(f"check_sass: {waits} kernels wait on the bank fill; {len(bad)} read the bank before the wait")
    for fn, st in bad[:20]:
        print("  VIOLATION", fn, st)
    return 1 if bad or waits == 0 else 0


if __name__ == "__main__":
    sys.exit(main())
