"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python profiles/launch_list_summary.py <csv> [bench json]"""
import collections, csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = []
for r in rows[start:]:
    if len(r) > vi:
        try:
            seq.append((r[ki], float(r[vi].replace(",", "")) / 1000.0))
        except ValueError:
            pass
ours = [(k, v) for k, v in seq if "cvb::" in k]
agg = collections.OrderedDict()
for k, v in ours:
    agg.setdefault(k.split("(")[0], []).append(v)
print("launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-vae-step --no-other-configs`")
print("(ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES, not absolutes)\n")
for k, v in agg.items():
    print(f"  {len(v):3d} launches  mean {sum(v) / len(v):8.2f} us   {k}")
fwd = [v for k, v in ours if "clifford_fwd_kernel<11, 1, 1, 0, 1>" in k]
bnd = [v for k, v in ours if "bind_v3_kernel<11, 0, 0>" in k]
if fwd and bnd:
    f, b = sum(fwd) / len(fwd), sum(bnd) / len(bnd)
    print(f"\nheadline step = clifford_fwd_kernel<11,PsRng,rowk,lean> + bind_v3_kernel<11,Mul,direct>:")
    print(f"  shares under ncu: sampler {100 * f / (f + b):.1f} %  bind {100 * b / (f + b):.1f} %")
    if len(sys.argv) > 2:
        d = json.load(open(sys.argv[2]))["kernels"]
        f2, b2 = d["rsample_kl"]["ms"], d["bind"]["ms"]
        print(f"  bench.py's live CUDA-event split (the bench JSON given: 20 timed steps of the default command on the same build): rsample_kl {f2:.4f} ms / bind {b2:.4f} ms = "
              f"{100 * f2 / (f2 + b2):.1f} % / {100 * b2 / (f2 + b2):.1f} %")
print(f"\nother launches in the list ({len(seq) - len(ours)}): torch's fills of the synthetic inputs and the e2e leg's device copies")
