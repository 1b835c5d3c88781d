cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/gpu/debug_symm.py > gpurun_out/debug_symm.out 2>&1
echo "rc=$?" >> gpurun_out/debug_symm.out
tail -30 gpurun_out/debug_symm.out
cat gpurun_out/debug_symm_r0.log | tail -30
