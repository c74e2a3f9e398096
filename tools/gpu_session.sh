#!/bin/bash
# One gpurun call = one session on the B200 box.  Every stage writes its log under gpurun_out/ and never aborts the
# following stages.  Usage: bash tools/gpu_session.sh stage1 stage2 ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc >> gpurun_out/gpu_info.txt
for stage in "$@"; do
  echo "=== stage $stage $(date +%T)"
  case "$stage" in
    small)    timeout 600 python tools/gpu_check.py --golden > gpurun_out/check_small.log 2>&1; echo "rc=$?" ;;
    big)      timeout 900 python tools/gpu_check.py --big > gpurun_out/check_big.log 2>&1; echo "rc=$?" ;;
    pytest)   timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" ;;
    pyteststream) G4R_TUNE_STREAM=1 timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu_stream.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu_stream.log ;;
    pytestall) timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" ;;
    smoke)    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" ;;
    memcheck) timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/memcheck.log 2>&1; echo "rc=$?" ;;
    racecheck) timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/racecheck.log 2>&1; echo "rc=$?" ;;
    initcheck) timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/initcheck.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/initcheck.log ;;
    memory)   timeout 600 python tools/retained_memory.py > gpurun_out/retained_memory.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/retained_memory.log | cut -c1-900 ;;
    benchmore) for wl in C2 C3sh3; do for impl in ours reference; do
                timeout 600 python bench.py --workload $wl --impl $impl --steps 100 --no-cpu-baseline > gpurun_out/bench_${wl}_${impl}.json 2> gpurun_out/bench_${wl}_${impl}.err; echo "$wl $impl rc=$?"; cut -c1-300 gpurun_out/bench_${wl}_${impl}.json
              done; done ;;
    bench)    timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json ;;
    benchref) timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches.log 2>&1; echo "rc=$?" ;;
    ncufull)  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:composite -s 4 -c 4 -o gpurun_out/prof_composite -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncufull.log 2>&1; echo "rc=$?" ;;
    ncuall)   timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'project_kernel|tile_scan|scatter_kernel|tile_sort|composite|gaussian_backward' -s 21 -c 7 -o gpurun_out/prof_all -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncuall.log 2>&1; echo "rc=$?"
              # summarise on the box so that a bench stage later in this session reports this capture's DRAM traffic
              python tools/ncu_summary.py gpurun_out/prof_all.ncu-rep "${G4R_TAG:-r01_v7}" && cp profiles/ncu_traffic.json profiles/${G4R_TAG:-r01_v7}_ncu_summary.md gpurun_out/ ;;
    sharded2) for wl in small X2 X4; do timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py --workload $wl > gpurun_out/sharded_$wl.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/sharded_$wl.log | cut -c1-1800; done ;;
    sharded2py) for wl in small X4; do G4R_SHARD_NATIVE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 tools/sharded_check.py --workload $wl --out gpurun_out/py > gpurun_out/sharded_py_$wl.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/sharded_py_$wl.log | cut -c1-1800; done ;;
    shardedN) n=${G4R_NGPU:-8}; for wl in ${G4R_SHARD_WL:-X4}; do timeout ${G4R_TIMEOUT:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 tools/sharded_check.py --workload $wl > gpurun_out/sharded_${wl}_x$n.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/sharded_${wl}_x$n.log | cut -c1-2500; done ;;
    benchN)   n=${G4R_NGPU:-8}; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $n --steps 100 > gpurun_out/bench_x$n.json 2> gpurun_out/bench_x$n.err; echo "rc=$?"; cat gpurun_out/bench_x$n.json | cut -c1-4000; tail -5 gpurun_out/bench_x$n.err ;;
    bench2)   timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --no-cpu-baseline > gpurun_out/bench_x2.json 2> gpurun_out/bench_x2.err; echo "rc=$?"; cat gpurun_out/bench_x2.json ;;
    shard8)   for n in 2 4 8; do
                timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n tools/sharded_check.py --workload C4 > gpurun_out/sharded_C4_x$n.log 2>&1; echo "sharded C4 x$n rc=$?"; tail -2 gpurun_out/sharded_C4_x$n.log | cut -c1-1500
              done ;;
    scale8)   for n in 2 4 8; do
                timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 100 --no-cpu-baseline > gpurun_out/bench_x$n.json 2> gpurun_out/bench_x$n.err; echo "bench x$n rc=$?"; cut -c1-200 gpurun_out/bench_x$n.json
                timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n tools/sharded_check.py --workload C4 > gpurun_out/sharded_C4_x$n.log 2>&1; echo "sharded C4 x$n rc=$?"; tail -2 gpurun_out/sharded_C4_x$n.log | cut -c1-400
              done ;;
    matrix)   timeout 900 python tools/perf_matrix.py > gpurun_out/perf_matrix.log 2>&1; echo "rc=$?"; cat gpurun_out/perf_matrix.log | cut -c1-400 ;;
    tune)     timeout 1200 python tools/tune_matrix.py > gpurun_out/tune_matrix.log 2>&1; echo "rc=$?"; cut -c1-300 gpurun_out/tune_matrix.log ;;
    hostline) timeout 600 python tools/host_timeline.py > gpurun_out/host_timeline.log 2>&1; echo "rc=$?"; cut -c1-1200 gpurun_out/host_timeline.log ;;
    determinism) timeout 900 python tools/determinism_check.py > gpurun_out/determinism.log 2>&1; echo "rc=$?"; cut -c1-2500 gpurun_out/determinism.log ;;
    repro)    timeout 900 python tools/repro_anomaly.py 16 > gpurun_out/repro_anomaly.log 2>&1; echo "rc=$?"; cut -c1-1500 gpurun_out/repro_anomaly.log ;;
    prelude)  timeout 600 python tools/prelude_bench.py > gpurun_out/prelude_bench.log 2>&1; echo "rc=$?"; cut -c1-700 gpurun_out/prelude_bench.log ;;
    knn)      timeout 900 python tools/knn_digests.py --write gpurun_out/knn_digests_ref.json > gpurun_out/knn_digests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/knn_digests.log | cut -c1-400
              [ -f gpurun_out/knn_digests_ref.json ] && [ ! -f tests/golden/knn_digests_ref.json ] && cp gpurun_out/knn_digests_ref.json tests/golden/
              timeout 900 python -m pytest tests/test_knn.py -m gpu -q --timeout 600 > gpurun_out/pytest_knn.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_knn.log ;;
    deform)   timeout 900 python -m pytest tests/test_deform.py -m gpu -q --timeout 600 > gpurun_out/pytest_deform.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_deform.log | cut -c1-300
              timeout 600 python tools/deform_bench.py > gpurun_out/deform_bench.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/deform_bench.log ;;
    knnsan)   timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/knn_digests.py K1_depthmap_20k K5_duplicates_60k > gpurun_out/knn_memcheck.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/knn_memcheck.log ;;
    digests)  timeout 900 python tools/digests.py --write gpurun_out/digests_ref.json > gpurun_out/digests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/digests.log ;;
    shapes)   for wl in C3map track; do for impl in ours reference; do
                timeout 600 python bench.py --workload $wl --impl $impl --steps 100 > gpurun_out/bench_${wl}_${impl}.json 2> gpurun_out/bench_${wl}_${impl}.err; echo "$wl $impl rc=$?"; cut -c1-600 gpurun_out/bench_${wl}_${impl}.json; tail -3 gpurun_out/bench_${wl}_${impl}.err
              done; done ;;
    *) echo "unknown stage $stage" ;;
  esac
done
for f in gpurun_out/check_small.log gpurun_out/pytest_gpu.log gpurun_out/check_big.log gpurun_out/smoke.log gpurun_out/memcheck.log; do
  [ -f "$f" ] && { echo "----- tail $f"; tail -n 25 "$f"; }
done
exit 0
