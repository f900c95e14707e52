"""Per-kernel SASS opcode counts of the built library (the evidence file under profiles/): python tools/sass_evidence.py"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "clifford-vae_b200", "clifford_b200", "libclifford_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
d = collections.defaultdict(collections.Counter)
fn = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        d[fn][m.group(1)] += 1
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip().split("(")[0]
want = ["clifford_fwd_kernelILi11ELi1ELb1ELb0ELb1E", "clifford_fwd_kernelILi11ELi1ELb1ELb1ELb0E", "clifford_fwd_kernelILi11ELi0ELb1ELb0ELb0E",
        "bind_v3_kernelILi9ELi0ELi0E", "bind_v3_kernelILi11ELi0ELi0E", "bind_v3_kernelILi12ELi0ELi2E", "bind_v3_kernelILi13ELi0ELi2E",
        "bind_pad_kernelILi10ELi0E", "clifford_bwd_kernelILi11ELb1ELb1E", "clifford_log_prob_kernelILi11ELb1ELb1E",
        "depth_chain_kernelILi12E"]
print("# r02 SASS evidence: cuobjdump -sass of clifford-vae_b200/clifford_b200/libclifford_b200.so (the library the round-2 bench runs)")
print("# packed fp32 = FADD2 / FMUL2 / FFMA2 (complex add = 1, complex multiply = 2 instructions); UBLKCP = cp.async.bulk (TMA, 1-D")
print("# form) staging of latent rows; SYNCS = mbarrier; no tensor-core opcode in the library (profiles/r02_dft_as_gemm.txt)")
for w in want:
    for f, ops in d.items():
        if w in f:
            print(f"{demangle(f)[:86]:88s} FADD2 {ops['FADD2']:4d} FMUL2 {ops['FMUL2']:4d} FFMA2 {ops['FFMA2']:4d} | scalar FADD {ops['FADD']:4d} FMUL {ops['FMUL']:4d} "
                  f"FFMA {ops['FFMA']:4d} | UBLKCP {ops['UBLKCP']:2d} SYNCS {ops['SYNCS']:2d} | {sum(ops.values())} instr")
tot = collections.Counter()
for ops in d.values():
    tot.update(ops)
print()
print("whole library (%d kernels):" % len(d), {k: tot[k] for k in ("FADD2", "FMUL2", "FFMA2", "UBLKCP", "SYNCS", "HMMA", "UTCHMMA", "LDTM", "UTMALDG")})
mb = os.path.join(ROOT, "tools", "microbench", "dft16_tcgen05.bin")
if os.path.exists(mb):
    o2 = subprocess.run(["cuobjdump", "-sass", mb], capture_output=True, text=True).stdout
    c = collections.Counter(re.findall(r"\b(UTCHMMA|LDTM|UTCBAR|UTCATOMSWS)\b", o2))
    print("tools/microbench/dft16_tcgen05.bin (the measured tensor-core prototype, not part of the library):", dict(c))
