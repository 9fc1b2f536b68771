run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', 'ms/step', round(d['ms_per_step'],3), 'value %.3e'%d['value'])"; }
run 29601 default
export NCCL_MAX_NCHANNELS=1 NCCL_MIN_NCHANNELS=1; run 29602 nch1; unset NCCL_MAX_NCHANNELS NCCL_MIN_NCHANNELS
export NCCL_PROTO=LL; run 29603 protoLL; unset NCCL_PROTO
export NCCL_PROTO=LL NCCL_MAX_NCHANNELS=2; run 29604 LL_nch2; unset NCCL_PROTO NCCL_MAX_NCHANNELS
export NCCL_NVLS_ENABLE=0; run 29605 nvls0; unset NCCL_NVLS_ENABLE
