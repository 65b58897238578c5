#!/bin/bash
# The one gpurun driver script.  Usage (from the repo root, on the GPU box):
#   tools/gpu.sh tests [pytest args]            GPU parity tests
#   tools/gpu.sh bench <tag> [bench.py args]    one bench line -> gpurun_out/bench_<tag>.json (+ digest)
#   tools/gpu.sh launches <tag> [bench.py args] ncu launch list of a short bench command
#   tools/gpu.sh ncu <tag> <kernel regex> [bench.py args]   ncu --set full of one launch (after the warm-up launches)
#   tools/gpu.sh metrics <tag> <kernel regex> [bench.py args]   DRAM bytes / duration / pipe utilisation of one launch (cheap)
#   tools/gpu.sh sanitize <tool> <tag> <python script> [args]   compute-sanitizer (memcheck | racecheck | synccheck)
#   tools/gpu.sh py <script> [args] | sh <shell command>        anything else, with PYTHONPATH set to the repo
# Several commands can be chained with ' -- ' between them:  tools/gpu.sh tests -x -- bench c2 --steps 20
mkdir -p gpurun_out
run_one() {
  cmd=$1; shift
  case $cmd in
    tests)
      timeout 1500 python -m pytest tests -m gpu -q "$@" 2>&1 | tail -25 ;;
    bench)
      tag=$1; shift
      timeout 900 python bench.py "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
      echo "bench $tag rc=$?"; python tools/bench_short.py < gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err ;;
    torchrun)
      n=$1; tag=$2; shift 2
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@" \
          > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
      echo "torchrun $tag rc=$?"; python tools/bench_short.py < gpurun_out/bench_$tag.json; grep -v "^W\|^\[W\|^$" gpurun_out/bench_$tag.err | tail -5 ;;
    launches)
      tag=$1; shift
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$tag.csv \
          python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check --named none "$@" > gpurun_out/launches_$tag.log 2>&1
      echo "launch list $tag rc=$?" ;;
    ncu)
      tag=$1; rx=$2; shift 2
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o gpurun_out/prof_$tag \
          python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check --named none "$@" > gpurun_out/ncu_$tag.log 2>&1
      echo "ncu $tag rc=$?"; ls -la gpurun_out/prof_$tag.ncu-rep
      # the text summary is made on the box; the report itself only travels back while gpurun_out stays small (64 MiB limit)
      python tools/ncu_summary.py gpurun_out/prof_$tag.ncu-rep > gpurun_out/ncu_summary_$tag.txt 2>&1
      if [ "$(du -sm gpurun_out | cut -f1)" -gt 40 ]; then rm -f gpurun_out/prof_$tag.ncu-rep; echo "(report dropped, summary kept)"; fi ;;
    metrics)
      # a few raw metrics of one launch (1-2 passes: cheap even for second-long kernels): DRAM bytes, duration, pipe utilisation
      tag=$1; rx=$2; shift 2
      timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
          --clock-control none -k regex:$rx -s 3 -c 1 --csv --log-file gpurun_out/metrics_$tag.csv \
          python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check --named none "$@" > gpurun_out/metrics_$tag.log 2>&1
      echo "metrics $tag rc=$?"; grep -v "^==" gpurun_out/metrics_$tag.csv | cut -d, -f5,12- | tail -9 ;;
    sanitize)
      tool=$1; tag=$2; shift 2
      timeout 1500 env PYTHONPATH=$PWD compute-sanitizer --tool $tool --print-limit 8 python "$@" > gpurun_out/sanitize_${tool}_$tag.txt 2>&1
      echo "sanitize $tool $tag rc=$?"
      grep -v "^=========     Host Frame\|^=========         in \|^=========                in" gpurun_out/sanitize_${tool}_$tag.txt | tail -30 ;;
    py)
      timeout 1500 env PYTHONPATH=$PWD python "$@" ;;
    sh)
      timeout 1500 env PYTHONPATH=$PWD bash -c "$*" ;;
    *) echo "unknown command $cmd"; return 2 ;;
  esac
}
args=()
for a in "$@"; do
  if [ "$a" = "--" ]; then run_one "${args[@]}"; args=(); else args+=("$a"); fi
done
[ ${#args[@]} -gt 0 ] && run_one "${args[@]}"
