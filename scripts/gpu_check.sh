# GPU-box check: smoke, parity tests, bench, ncu launch list + full captures of the headline kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
OSB_SP_CONV=ffma timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -60
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" = "full" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-solve --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
tail -3 gpurun_out/bench_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:db_scan -s 2 -c 4 -o gpurun_out/prof_dbscan -f python bench.py --steps 2 --warmup 1 --no-solve --no-cpu-baseline > gpurun_out/ncu_dbscan.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_umma_kernel<\(int\)64, \(bool\)1,'  -s 3 -c 3 -o gpurun_out/prof_conv -f python bench.py --steps 2 --warmup 1 --no-solve --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:graph_solve -s 1 -c 1 -o gpurun_out/prof_solve -f python -c "
import sys; sys.path.insert(0, '.')
from omniswarm_b200 import host, synth
g = synth.pose_graph_c5(0); s = host.PoseGraphSolver(2048, 12288)
for _ in range(2): p, m = s.solve(g)
print(m.solve_ms, m.pcg_iterations)
" > gpurun_out/ncu_solve.log 2>&1
fi
ls -la gpurun_out
