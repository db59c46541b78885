python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2f_gpu_tests.log; cat gpurun_out/r2f_gpu_tests.log
timeout 600 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_racecheck.log 2>&1; tail -3 gpurun_out/r2f_racecheck.log
python bench.py > gpurun_out/r2f_c3.json 2> gpurun_out/r2f_c3.err; echo c3 rc=$?; tail -c 600 gpurun_out/r2f_c3.json | head -c 300; echo
python bench.py --workload c4 > gpurun_out/r2f_c4.json 2> gpurun_out/r2f_c4.err; echo c4 rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2f_ncu_bench.log 2>&1; echo launches rc=$?
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:"k_select_lookup|k_score_cta|k_score_warp|k_s1_finish|k_wm|k_cov_eval|k_expand" --launch-count 8 -f -o gpurun_out/r2f_top python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2f_ncu_full.log 2>&1; echo full rc=$?
python -c "
import json
for f in ('gpurun_out/r2f_c3.json','gpurun_out/r2f_c4.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], d['phases_ms_per_step'], d['roofline']['frac'], d['parity'], d['clocks'])"
