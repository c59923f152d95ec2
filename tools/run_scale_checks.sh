# usage: bash tools/run_scale_checks.sh N TAG   (inside gpurun --gpus N)
N=$1; TAG=$2
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > gpurun_out/${TAG}_$2.json 2> gpurun_out/${TAG}_$2.err; tail -1 gpurun_out/${TAG}_$2.err; }
run 29521 n${N}_peer_ops --steps 20 --warmup 3 --ops
run 29522 n${N}_nccl --steps 20 --warmup 3 --no-sweep --exchange nccl
python - <<PY
import json
for k in ("n${N}_peer_ops","n${N}_nccl"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json"%k))
        print(k, d["ms_per_step"], d["value"], d["chain"]["frac_of_aggregate_peak"], d["e2e"]["value"], d["result"], d.get("ops"))
    except Exception as e: print(k, "failed", e)
PY
