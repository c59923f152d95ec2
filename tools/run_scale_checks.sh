# usage: bash tools/run_scale_checks.sh N TAG   (inside gpurun --gpus N)
# The driver's own command line at N GPUs (default flags: chain + exchange + sharded hash join),
# then the same with the NCCL exchange for comparison.
N=$1; TAG=$2
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > gpurun_out/${TAG}_$2.json 2> gpurun_out/${TAG}_$2.err; tail -1 gpurun_out/${TAG}_$2.err; }
run 29521 n${N}_peer --steps 20 --warmup 3
run 29522 n${N}_nccl --steps 20 --warmup 3 --no-sweep --no-join --exchange nccl
python - <<PY
import json
for k in ("n${N}_peer","n${N}_nccl"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json"%k))
        print(k, d["ms_per_step"], d["value"], d["chain"]["frac_of_aggregate_peak"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["result"])
        print(json.dumps(d.get("hash_join"))[:1500])
    except Exception as e: print(k, "failed", e)
PY
