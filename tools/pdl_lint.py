"""SASS lint for programmatic dependent launch: in every kernel that executes griddepcontrol.wait (SASS:
ACQBULK) no global load may be scheduled ahead of it, unless the kernel is listed in ALLOWED (loads of data
that is constant across the kernel chain).  ptxas treats ld.global.nc (`const T* __restrict__`) as
invariant and WILL hoist it above the wait -- reading the predecessor's output before it is written.

    python tools/pdl_lint.py [path/to/librg_b200.so]      exit code 1 on violations
"""
import os
import re
import subprocess
import sys

ALLOWED = {}     # kernel-name substring -> why its early loads are safe (none at present)


def scan(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels, name, ins = {}, None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                kernels[name] = ins
            name, ins = m.group(1), []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and name:
            ins.append(m.group(1).strip())
    if name:
        kernels[name] = ins
    report = {}
    for k, ins in kernels.items():
        waits = [i for i, s in enumerate(ins) if "ACQBULK" in s]
        if not waits:
            continue
        early = [s for s in ins[:waits[0]] if re.search(r"\b(LDG|LD\.E|LD\b)", s)]
        report[k] = early
    return report


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                            "rag_gesture_b200", "librg_b200.so")
    bad = 0
    for k, early in sorted(scan(lib).items()):
        short = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        short = short.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][-60:]
        allowed = next((why for a, why in ALLOWED.items() if a in k), None)
        if early and not allowed:
            bad += 1
            print(f"VIOLATION {short}: {len(early)} global loads above griddepcontrol.wait, e.g. {early[0]}")
        else:
            print(f"ok        {short}: {len(early)} early loads" + (f" (allowed: {allowed})" if early else ""))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
