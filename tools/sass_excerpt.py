"""SASS evidence for the pair-force kernel: mnemonic counts of an instantiation and the excerpt between two markers.
usage: python tools/sass_excerpt.py > profiles/r02_sass_force_tile4.txt   (reads cellflow_b200/lib/libcellflow_b200.so)"""
import collections, re, subprocess, sys

LIB = "cellflow_b200/lib/libcellflow_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((m.group(1), re.sub(r"\s+", " ", m.group(2)).strip()))


def find(tag):
    return next(v for k, v in funcs.items() if tag in k)


def counts(ins):
    c = collections.Counter()
    for _, t in ins:
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        c[op.split(".")[0]] += 1
    keys = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "MUFU", "FSETP", "FMNMX", "FMNMX3", "VOTE", "LDS", "STS",
            "LDG", "STG", "UBLKCP", "SYNCS", "POPC", "BREV", "FLO", "REDUX", "SHFL", "BRA", "LOP3", "IADD3", "IMAD"]
    return "  ".join(f"{k}:{c[k]}" for k in keys)


def excerpt(ins, first, last, occurrence=0):
    """instructions from the `occurrence`-th match of regex `first` up to the next match of `last`"""
    starts = [i for i, (_, t) in enumerate(ins) if re.search(first, t)]
    a = starts[occurrence]
    b = next(i for i in range(a + 1, len(ins)) if re.search(last, ins[i][1]))
    return "\n".join(f"    /*{ad}*/ {t}" for ad, t in ins[a:b + 1])


print("# SASS of the pair-force kernel, round 2 (cuobjdump -sass cellflow_b200/lib/libcellflow_b200.so, nvcc 12.9, sm_100a)\n")
print("Produced by tools/sass_excerpt.py.  Mnemonic counts over whole instantiations, then excerpts of the default kernel\n"
      "(compacted staging, STAGE 4): the chunk path (4 x LDG.E.128 with clamped indices -> quad box -> box prefilter vote ->\n"
      "compacted 3 x STS.128) and the block loop (3 x LDS.128 at a running address, 12 packed FADD2/FMUL2/FFMA2, min, vote, and\n"
      "the force terms of a live block: 8 MUFU, packed FMUL2/FFMA2, 4 FSETP + 12 predicated FFMA + 4 predicated adds); last\n"
      "the bulk-copy variant's request (UBLKCP) for the A/B of profiles/r02_staging_ab.md.\n")
m1 = find("force_tile4_kernelILi1ELb0ELi4ELi1E")
m0 = find("force_tile4_kernelILi0ELb0ELi4ELi2E")
old = find("force_tile4_kernelILi1ELb0ELi0ELi1E")
blk = find("force_tile4_kernelILi1ELb0ELi1ELi1E")
print(f"## per-type radii, 1 layer, compacted staging (default) <1,false,4,1>: {len(m1)} instructions\n  {counts(m1)}\n")
print(f"## uniform radius, 2 layers, compacted staging (default) <0,false,4,2>: {len(m0)} instructions\n  {counts(m0)}\n")
print(f"## per-type radii, every quad stored (option t4_stage=0) <1,false,0,1>: {len(old)} instructions\n  {counts(old)}\n")
print(f"## per-type radii, cp.async.bulk staging (option t4_stage=1) <1,false,1,1>: {len(blk)} instructions\n  {counts(blk)}\n")
print("### <1,false,4,1> chunk path: loads, quad box, prefilter vote, compacted store\n")
print(excerpt(m1, r"LDG\.E\.128 ", r"STS\.128 .*0x2000\]"))
print("\n### <1,false,4,1> block loop (no-wrap path): exact test of one (layer, quad) block, vote, force terms, loop test\n")
print(excerpt(m1, r"LDS\.128 R\d+, \[R\d+\+0x1000\]", r"@P\d BRA", occurrence=1))
print("\n### <1,false,1,1> request of the next chunk: mbarrier arm + three bulk copies\n")
print(excerpt(blk, r"SYNCS\.ARRIVE\.TRANS64|ELECT", r"UBLKCP", occurrence=0))
