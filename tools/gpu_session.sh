#!/bin/bash
# Development helper: what one `gpurun -- bash tools/gpu_session.sh [steps...]` call runs.  EVERY step has its own short timeout: a
# kernel that hangs must cost a minute, not the round's GPU budget (it did once: profiles/r01_SUMMARY.md, "two lanes per G2 bucket").
# steps: solver tests bench22 bench26 ref26 ncu_launches ncu_poseidon microbench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
steps="${@:-solver tests bench22}"
for st in $steps; do
  case $st in
    solver)  timeout 600 python -m pytest tests/test_gpu_solver.py -x -q > gpurun_out/s_solver.log 2>&1; rc=$?; echo "solver tests rc=$rc"; tail -15 gpurun_out/s_solver.log
             [ $rc -ne 0 ] && exit $rc ;;
    tests)   timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_solver.py > gpurun_out/s_tests.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -6 gpurun_out/s_tests.log
             [ $rc -ne 0 ] && exit $rc ;;
    bench22) timeout 600 python bench.py --log-n 22 --cpu-log-n 18 > gpurun_out/s_bench22.json 2> gpurun_out/s_bench22.err; echo "bench22 rc=$?"; tail -c 1500 gpurun_out/s_bench22.json; tail -5 gpurun_out/s_bench22.err ;;
    bench26) timeout 1200 python bench.py --no-cpu > gpurun_out/s_bench26.json 2> gpurun_out/s_bench26.err; echo "bench26 rc=$?"; tail -c 2500 gpurun_out/s_bench26.json; tail -5 gpurun_out/s_bench26.err ;;
    bench26cpu) timeout 1500 python bench.py > gpurun_out/s_bench26.json 2> gpurun_out/s_bench26.err; echo "bench26 rc=$?"; tail -c 2500 gpurun_out/s_bench26.json; tail -5 gpurun_out/s_bench26.err ;;
    ref26)   timeout 1700 python bench.py --impl reference > gpurun_out/s_ref26.json 2> gpurun_out/s_ref26.err; echo "ref26 rc=$?"; tail -c 1500 gpurun_out/s_ref26.json; tail -5 gpurun_out/s_ref26.err ;;
    ref24)   timeout 900 python bench.py --impl reference --log-n 24 > gpurun_out/s_ref24.json 2> gpurun_out/s_ref24.err; echo "ref24 rc=$?"; tail -c 1500 gpurun_out/s_ref24.json; tail -5 gpurun_out/s_ref24.err ;;
    ncu_launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k 'regex:^(?!k_solve_wide|k_solve_div).*' -c 4000 --csv --log-file gpurun_out/launches_bench22.csv python bench.py --log-n 22 --steps 1 --warmup 1 --provers 1 --no-cpu --no-e2e --no-parity > gpurun_out/s_ncu_l.log 2>&1; echo "ncu launches rc=$?" ;;
    solverbench22) timeout 600 python tools/solver_bench.py 22 > gpurun_out/s_solverbench22.log 2>&1; echo "solverbench22 rc=$?"; tail -8 gpurun_out/s_solverbench22.log | cut -c1-600 ;;
    solverbench26) timeout 900 python tools/solver_bench.py 26 96:512:9 96:512:0 > gpurun_out/s_solverbench26.log 2>&1; echo "solverbench26 rc=$?"; tail -4 gpurun_out/s_solverbench26.log | cut -c1-600 ;;
    sharded) timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/s_sharded.log 2>&1; rc=$?; echo "sharded tests rc=$rc"; tail -25 gpurun_out/s_sharded.log
             [ $rc -ne 0 ] && exit $rc ;;
    bench22n) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NGPU --log-n 22 --no-cpu > gpurun_out/s_bench22_n$NGPU.json 2> gpurun_out/s_bench22_n$NGPU.err; echo "bench22n rc=$?"; tail -c 3000 gpurun_out/s_bench22_n$NGPU.json; tail -5 gpurun_out/s_bench22_n$NGPU.err ;;
    bench26n) timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NGPU --no-cpu --no-parity > gpurun_out/s_bench26_n$NGPU.json 2> gpurun_out/s_bench26_n$NGPU.err; echo "bench26n rc=$?"; tail -c 3000 gpurun_out/s_bench26_n$NGPU.json; tail -5 gpurun_out/s_bench26_n$NGPU.err ;;
    solvertrace26) ZKPOR_SOLVE_TRACE=gpurun_out/solve_trace26.csv timeout 900 python tools/solver_bench.py 26 96:512:9 > gpurun_out/s_solvertrace26.log 2>&1; echo "solvertrace26 rc=$?"; tail -40 gpurun_out/s_solvertrace26.log | cut -c1-400 ;;
    inflight26) timeout 900 python tools/inflight_bench.py 26 2 3 > gpurun_out/s_inflight26.log 2>&1; echo "inflight26 rc=$?"; tail -6 gpurun_out/s_inflight26.log | cut -c1-400 ;;
    inflight22) timeout 600 python tools/inflight_bench.py 22 2 3 > gpurun_out/s_inflight22.log 2>&1; echo "inflight22 rc=$?"; tail -6 gpurun_out/s_inflight22.log | cut -c1-400 ;;
    witness) timeout 900 python bench.py --workload witness > gpurun_out/s_witness.json 2> gpurun_out/s_witness.err; echo "witness rc=$?"; tail -c 2500 gpurun_out/s_witness.json; tail -5 gpurun_out/s_witness.err ;;
    witness_small) timeout 600 python bench.py --workload witness --accounts 400000 > gpurun_out/s_witness_small.json 2> gpurun_out/s_witness_small.err; echo "witness_small rc=$?"; tail -c 2500 gpurun_out/s_witness_small.json; tail -5 gpurun_out/s_witness_small.err ;;
    ncu_poseidon) timeout 900 ncu --set full --import-source on --clock-control none -k regex:'k_account_leaves_tpa|k_merkle_level|k_cex_commitments' --launch-skip 2 -c 6 -f -o gpurun_out/r02_poseidon python bench.py --workload witness --accounts 600000 > gpurun_out/s_ncu_poseidon.log 2>&1; echo "ncu poseidon rc=$?"; tail -3 gpurun_out/s_ncu_poseidon.log | cut -c1-300 ;;
    ncu_accumulate) timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_accumulate -c 4 -f -o gpurun_out/r02_accumulate python tools/ncu_traffic.py run 24 > gpurun_out/s_ncu_acc.log 2>&1; echo "ncu accumulate rc=$?"; tail -3 gpurun_out/s_ncu_acc.log | cut -c1-300 ;;
    ncu_narrow) timeout 900 ncu --set full --import-source on --clock-control none -k k_solve_narrow --launch-skip 6 -c 3 -f -o gpurun_out/r02_narrow python tools/solver_bench.py 22 96:512:9 > gpurun_out/s_ncu_narrow.log 2>&1; echo "ncu narrow rc=$?"; tail -3 gpurun_out/s_ncu_narrow.log | cut -c1-300 ;;
    *) echo "unknown step $st" ;;
  esac
done
