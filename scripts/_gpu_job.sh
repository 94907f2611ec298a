python -m pytest tests -m gpu -q > gpurun_out/r2q_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2q_tests.log
tail -8 gpurun_out/r2q_tests.log
HW_N=16384 HW_T=40 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/nccl_collector_check.py > gpurun_out/r2q_nccl_collector.json 2> gpurun_out/r2q_nccl_collector.err; echo "nccl exit $?"
cat gpurun_out/r2q_nccl_collector.json; tail -5 gpurun_out/r2q_nccl_collector.err
