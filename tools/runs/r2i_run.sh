TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
python tools/p2p_ce_bw.py 2>&1 | tail -1
NCCL_MAX_NCHANNELS=64 NCCL_MIN_NCHANNELS=64 NCCL_MIN_P2P_NCHANNELS=64 NCCL_MAX_P2P_NCHANNELS=64 $TR tools/p2p_bw.py 2>/dev/null | tail -1
NCCL_P2P_USE_CUDA_MEMCPY=1 $TR tools/p2p_bw.py 2>/dev/null | tail -1
NCCL_NTHREADS=512 NCCL_MAX_P2P_NCHANNELS=32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_BUFFSIZE=16777216 $TR tools/p2p_bw.py 2>/dev/null | tail -1
