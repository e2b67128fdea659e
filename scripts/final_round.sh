# Round-end verification on one GPU: the whole GPU test suite, smoke(), every bench line, the in-graph cost tables.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_pytest_f.txt 2>&1; tail -3 gpurun_out/r2_pytest_f.txt
grep -h "^\[parity\|^\[ensemble\|^\[paired\|^\[trajectory\|^\[mega" gpurun_out/r2_pytest_f.txt > gpurun_out/r2_parity_errors.txt
grep -o "\[parity[^]]*\].*\|\[ensemble[^]]*\].*\|\[paired[^]]*\].*\|\[trajectory[^]]*\].*\|\[mega[^]]*\].*" gpurun_out/r2_pytest_f.txt > gpurun_out/r2_parity_errors.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r2_bench_final_den.json 2> gpurun_out/r2_bench_final_den.err; tail -c 600 gpurun_out/r2_bench_final_den.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_final_ref.json 2> gpurun_out/r2_bench_final_ref.err
for c in sr inp ct; do timeout 600 python bench.py --config $c > gpurun_out/r2_bench_final_$c.json 2> gpurun_out/r2_bench_final_$c.err; done
python - <<'PY'
import json
for c in ("den", "ref", "sr", "inp", "ct"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2_bench_final_{c}.json") if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        print(c, "%.2f %s  %.3f ms  e2e %.2f  roofline %s %.3g/%.4g = %.3f" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], r.get("kernel"), r.get("achieved", 0), r.get("peak", 1), r.get("frac", 0)))
    except Exception as e:
        print(c, "FAILED", e)
PY
for mc in 8 1; do python scripts/plan_link_cost.py tf32 $mc > gpurun_out/r2_plan_link_mc$mc.txt 2>&1; python scripts/link_cost.py $mc > gpurun_out/r2_link_cost_s$mc.txt 2>&1; done
tail -16 gpurun_out/r2_plan_link_mc8.txt
