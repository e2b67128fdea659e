"""Summarise an ncu CSV launch list (metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum) of
`bench.py --steps 2 --warmup 3 --no-cpu` into per-kernel time/traffic of ONE graph-replayed step and map CUDA kernels to
the C-ABI families bench.py reports.  usage: python scripts/traffic_from_ncu.py launches.csv out_prefix"""
import collections
import csv
import json
import sys

src, out = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
iID, iK, iM, iU, iV = (hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
launch = collections.OrderedDict()
for row in r:
    if len(row) <= iV:
        continue
    d = launch.setdefault(int(row[iID]), {"kernel": row[iK].split("(")[0].replace("void ", "")})
    v = float(row[iV].replace(",", ""))
    u = row[iU]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    d[row[iM]] = v * scale
items = list(launch.values())
# a step starts with the accumulator fill that precedes the input-jitter kernel (other k_fill launches zero dT buffers mid-step)
starts = [i - 1 for i, d in enumerate(items) if "k_input_jitter_pad" in d["kernel"] and i > 0]
step = items[starts[0]:starts[1]] if len(starts) >= 2 else items
fam_of = [("k_conv_halo", None), ("k_conv_pointwise", None), ("k_wgrad_alias", "mfvi_conv2d_wgrad"), ("k_wgrad_tc", "mfvi_conv2d_wgrad"),
          ("k_conv_wgrad", "mfvi_conv2d_wgrad"), ("k_bias_grad", "mfvi_conv2d_wgrad"), ("k_pad_act_bwd", "mfvi_pad_act_bwd"),
          ("k_bn_bwd_apply", "mfvi_bn_bwd_apply"), ("k_bn_act_pad_fwd", "mfvi_bn_act_pad_fwd"), ("k_cat_up_fwd", "mfvi_cat_up_fwd"),
          ("k_cat_bwd", "mfvi_cat_up_bwd"), ("k_kl_reparam", "mfvi_kl_reparam_fwd_bwd"), ("k_sample_weights", "mfvi_sample_weights"),
          ("k_adamw", "mfvi_adamw_step"), ("k_nll", "mfvi_gauss_nll_fwd_bwd"), ("k_fill", "mfvi_fill_f32"),
          ("k_input_jitter", "mfvi_input_jitter_pad")]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
fam = collections.defaultdict(float)
seen_nll = False
for d in step:
    k = d["kernel"]
    t = d.get("gpu__time_duration.sum", 0.0)
    b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a = agg[k]
    a[0] += 1; a[1] += t; a[2] += b
    if "k_nll" in k or "k_mse" in k:
        seen_nll = True
    f = None
    for pat, name in fam_of:
        if pat in k:
            f = name
            break
    if f is None and ("k_conv_halo" in k or "k_conv_tc" in k or "k_conv_igemm" in k or "k_conv_pointwise" in k):
        # forward launches come before the loss kernel; in the backward, wgrad is recognised above, the rest is dgrad
        f = "mfvi_conv2d_dgrad" if seen_nll else "mfvi_conv2d_fwd"
    if f:
        fam[f] += b
tot = sum(a[1] for a in agg.values())
with open(out + "_summary.txt", "w") as f:
    f.write(f"# one graph-replayed step: {len(step)} launches, {tot:.1f} us of kernel time (ncu, serialised, cold caches: compare shares)\n")
    f.write(f"{'kernel':34s} {'n':>4s} {'us':>9s} {'share':>6s} {'dram MB':>9s} {'GB/s':>7s}\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:34s} {a[0]:4d} {a[1]:9.1f} {a[1]/tot:6.3f} {a[2]/1e6:9.1f} {a[2]/max(a[1],1e-9)/1e3:7.0f}\n")
json.dump({"source": src, "bytes_per_step": dict(fam)}, open(out + "_dram_traffic.json", "w"), indent=1)
print(open(out + "_summary.txt").read())
