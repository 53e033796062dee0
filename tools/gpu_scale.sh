#!/bin/bash
# multi-GPU evidence (run under `gpurun --gpus N`): the frame / rollout / clip workloads of bench.py at N ranks (torchrun, NCCL)
# N defaults to the number of visible GPUs
N=${N:-$(nvidia-smi -L | wc -l)}
mkdir -p gpurun_out
run() {  # name, extra args
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 $2 > gpurun_out/r02_scale_$1_n$N.json 2> gpurun_out/r02_scale_$1_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $N $2 > gpurun_out/r02_scale_$1_n$N.json 2> gpurun_out/r02_scale_$1_n$N.err
  fi
  echo "$1 N=$N rc=$?"; tail -c 400 gpurun_out/r02_scale_$1_n$N.err | tail -3
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_scale_$1_n$N.json").read().strip().splitlines()[-1])
    print("$1", "n_gpus", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["config"].get("batch_per_gpu"), d.get("clocks"))
except Exception as e:
    print("$1 parse failed", e)
PY
}
run frame "--steps ${STEPS:-50} --warmup 3 --no-cpu-baseline"
run rollout "--workload rollout --steps ${RSTEPS:-20} --warmup 3 ${ROLLOUT_ARGS:-}"
if [ "${CLIP:-1}" = "1" ]; then run clip "--workload clip --steps ${RSTEPS:-20} --warmup 3"; fi
