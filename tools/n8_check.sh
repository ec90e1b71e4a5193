# one 8-GPU box: peer all-reduce exactness + bandwidth at 4 and 8 ranks, then the LM step at N=8 with both backends
export P2P_BENCH_CTAS=148,1184
for n in 4 8; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n tools/p2p_bench.py 2>&1 | grep "^p2p\|^nccl\|Error\|error" | sed "s/^/W=$n  /"
done
for v in A=1 SK_DP_COMM=nccl; do
  env $v timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 --skip-hubert --no-cpu-baseline 2>gpurun_out/n8_$v.err | tail -1 > gpurun_out/n8_$v.json
  python -c "import json,sys; d=json.loads(open('gpurun_out/n8_$v.json').read()); print('N=8 $v', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), {k: round(x,2) for k,x in d['roofline']['breakdown_ms'].items()}, d['clocks']['sm_mhz'], d['config'].get('dp_comm'), d.get('final_loss'))" || tail -5 gpurun_out/n8_$v.err
done
