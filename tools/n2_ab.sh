# same-box comparison of the LM step on one GPU vs two (run under `gpurun --gpus 2`): ms/step of both timed legs
python bench.py --steps 20 --warmup 5 --skip-hubert --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])"
for v in default SK_NO_OVERLAP=1 SK_DP_COMM=nccl; do
  env $( [ "$v" = default ] || echo $v ) python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --skip-hubert --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=2 $v', d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])"
done
