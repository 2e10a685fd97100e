"""Per-kernel SASS mnemonic counts of the built library -> profiles/sass_summary.txt (needs only cuobjdump, no GPU).

    python scripts/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import re
import subprocess
import sys

LIB = "curvature_b200/libcurvature_b200.so"
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMACCTL", "UBLKPF", "SYNCS", "HMMA", "FFMA", "DFMA", "ATOM/RED"]
PAT = {c: re.compile(r"\b" + c) for c in COLS if c != "ATOM/RED"}
PAT["ATOM/RED"] = re.compile(r"\b(ATOMG|REDG)\b")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return out[:len(names)]


def short(d):
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    d = re.sub(r"^void ", "", d)
    d = d.split("(")[0]
    return d.replace("crv::", "").replace("<(bool)1>", "<bf16>").replace("<(bool)0>", "<tf32>").replace("<true>", "<bf16>").replace("<false>", "<tf32>")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for c, p in PAT.items():
            if p.search(line):
                counts[cur][c] += 1
    names = list(counts)
    dem = [short(d) for d in demangle(names)]
    print(f"# SASS evidence of the Blackwell-native paths (`cuobjdump -sass {LIB}`; regenerate with `python scripts/sass_summary.py`)\n")
    print("| kernel | " + " | ".join(COLS) + " |")
    print("|" + "---|" * (len(COLS) + 1))
    tot = collections.Counter()
    for n, d in zip(names, dem):
        print(f"| {d} | " + " | ".join(str(counts[n][c]) for c in COLS) + " |")
        tot.update(counts[n])
    print("| **total** | " + " | ".join(str(tot[c]) for c in COLS) + " |")
    print("""
Legend: UTCHMMA = tcgen05.mma (kind::f16 / kind::tf32); UTCBAR = tcgen05.commit; LDTM = tcgen05.ld (TMEM -> registers);
UTMALDG = cp.async.bulk.tensor load (TMA); UTMASTG = cp.async.bulk.tensor store; UTMAREDG = cp.reduce.async.bulk.tensor
(reduce-add applied in L2); UTMACCTL = tensor-map prefetch (prefetch.tensormap); UBLKPF = cp.async.bulk.prefetch.L2; SYNCS = mbarrier;
HMMA = legacy mma.sync; FFMA / DFMA = fp32 / fp64 FMA; ATOM/RED = global atomics (ATOMG / REDG).

No HMMA anywhere: no kernel falls back to the legacy mma.sync path.  The tensor-core kernels (syrk_nhwc_kernel<bf16|tf32>,
syrk_tc_kernel, syrk_tc_tma_kernel, gemm_tc_kernel, gemm_chain_kernel) issue UTCHMMA from TMA-filled shared memory (UTMALDG)
into TMEM and drain it with LDTM; gemm_chain_kernel writes its results back through the TMA unit (UTMASTG / UTMAREDG).
syrk_nhwc_kernel also carries the (optional, off by default) TMA-store flush of its accumulator.  Global atomics
(ATOMG / REDG): the tile cursor and the per-row-tile dependency counters of gemm_chain_kernel, the fp32 merge of the
contraction splits in the CUDA-core SYRK kernels (syrk_simt*), and the debug trace slots (64-bit min / max).""")


if __name__ == "__main__":
    sys.exit(main())
