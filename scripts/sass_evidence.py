"""SASS evidence that the hot kernels are Blackwell-native (B200_PROFILING.md, "What proves a Blackwell-native kernel"):
counts, per kernel of grapes_b200/libgrapes_b200.so, of the mnemonics tcgen05.mma / tcgen05.ld / tcgen05.st / TMA / bulk copy /
mbarrier compile to, and of the legacy tensor path (HMMA, must be absent).  No GPU needed (cuobjdump reads the cubins).

    python scripts/sass_evidence.py            # prints the table (markdown)
    python scripts/sass_evidence.py --write    # also writes profiles/r02_sass_evidence.md
"""
import collections
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "grapes_b200", "libgrapes_b200.so")

# mnemonic -> what was written in the source
MNEMONICS = [("UTC*MMA", r"\bUTC[A-Z]*MMA\b", "tcgen05.mma"), ("LDTM", r"\bLDTM\b", "tcgen05.ld"), ("STTM", r"\bSTTM\b", "tcgen05.st"),
             ("UTCBAR", r"\bUTCBAR\b", "tcgen05.commit"), ("UTMALDG", r"\bUTMALDG\b", "cp.async.bulk.tensor (TMA tile load)"),
             ("UBLKCP", r"\bUBLKCP\b", "cp.async.bulk (TMA row copy)"), ("SYNCS", r"\bSYNCS\b", "mbarrier"),
             ("MEMBAR.*SYS", r"\bMEMBAR\.[A-Z]+\.SYS\b", "system-scope release/acquire (peer memory)"),
             ("HMMA", r"\bHMMA\b", "mma.sync / wmma (legacy, must be 0)")]


def cuobjdump():
    return shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


def kernel_families(mangled):
    """mangled names -> kernel name without template arguments / parameter list (the instantiations of one kernel are one row)."""
    res = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True, check=True).stdout.splitlines()
    fams = []
    for d in res:
        d = re.sub(r"^void\s+", "", d.strip())
        fams.append(re.split(r"[<(]", d, maxsplit=1)[0])
    return fams


def sass_counts(lib_path: str = LIB):
    """{kernel family: {"instances": n, mnemonic: total count over the instances}} and the arch set of the cubins."""
    out = subprocess.run([cuobjdump(), "-sass", lib_path], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    fams = collections.OrderedDict()
    parts = re.split(r"\n\s*Function : ", out)[1:]
    names = kernel_families([p.split("\n", 1)[0].strip() for p in parts])
    for part, fam in zip(parts, names):
        row = fams.setdefault(fam, collections.Counter())
        row["instances"] += 1
        for key, rx, _ in MNEMONICS:
            row[key] += len(re.findall(rx, part))
    return fams, archs


def table(fams, archs) -> str:
    keys = [k for k, _, _ in MNEMONICS]
    lines = ["# SASS evidence (round 2): `cuobjdump -sass grapes_b200/libgrapes_b200.so`", "",
             f"Cubin architectures: {', '.join(sorted(archs))} (an `sm_52` entry, if listed, is the EMPTY device-link stub `nvcc -shared` "
             f"adds: it holds no function).  {len(fams)} kernels ({sum(r['instances'] for r in fams.values())} "
             "instantiations).  Counts are static instruction counts summed over a kernel's template instantiations; only kernels "
             "with at least one of these mnemonics are listed.", "",
             "| source construct | SASS |", "|---|---|"]
    lines += [f"| {what} | `{k}` |" for k, _, what in MNEMONICS]
    lines += ["", "| kernel | instantiations | " + " | ".join(f"`{k}`" for k in keys) + " |", "|---|---|" + "---|" * len(keys)]
    for fam, row in fams.items():
        if any(row[k] for k in keys if k != "SYNCS"):
            lines.append(f"| `{fam}` | {row['instances']} | " + " | ".join(str(row[k]) for k in keys) + " |")
    total_h = sum(r["HMMA"] for r in fams.values())
    lines += ["", f"Legacy tensor-core instructions (`HMMA`) in the whole library: {total_h}.",
              "`k_step_tail` (the `<true>` instantiation) is the gradient scale + push all-reduce over NVLink peer memory + Adam: its "
              "system-scope fences bracket the posted stores into the peers' buffers and the acquire of this rank's own flags."]
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    fams, archs = sass_counts()
    md = table(fams, archs)
    print(md)
    if "--write" in sys.argv:
        with open(os.path.join(ROOT, "profiles", "r02_sass_evidence.md"), "w") as f:
            f.write(md)
