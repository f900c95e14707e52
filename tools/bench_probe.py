"""Short probe of the bench: headline value, kernel split and a few other_configs legs (python tools/bench_probe.py)."""
import json,sys,subprocess
out=subprocess.run([sys.executable,"bench.py","--steps","10","--warmup","3","--no-cpu-baseline","--no-vae-step"],capture_output=True,text=True).stdout
d=json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
o=d["other_configs"]
print("value %.4e"%d["value"], {k:round(v["ms"],5) for k,v in d["kernels"].items()})
for k in ("C3_rsample_backward_B4096_d2048","C3_log_prob_B4096_d2048","C5_depth_1_32_d8192","C1_clifford_rsample_kl_B128_d512"): print(k, round(o[k]["ms"],5), round(o[k]["frac"],4))
