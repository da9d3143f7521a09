#!/bin/bash
# Scaling run on one 8xB200 box (gpurun --gpus 8): bench.py at N = 1, 2, 4, 8 the way the driver launches it, the multi-GPU
# tests, the C-ABI group path (no torch) on the same frame, C5 full size at N = 8.
tag=${1:-r02}
python -m pytest tests/test_gpu_multi.py -x -q -s 2>&1 | tail -8 > gpurun_out/${tag}_multi_tests.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_c4_n1.json 2> gpurun_out/${tag}_c4_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_c4_n$n.json 2> gpurun_out/${tag}_c4_n$n.err
done
XML=/tmp/rt795_scenes/c4_full/c4.xml
: > gpurun_out/${tag}_group.txt
for devs in "0" "0 1" "0 1 2 3" "0 1 2 3 4 5 6 7"; do raytracer-795_b200/bin/group_render_test $XML $devs >> gpurun_out/${tag}_group.txt 2>&1; done
RT795_GROUP_NO_PEER=1 raytracer-795_b200/bin/group_render_test $XML 0 1 2 3 4 5 6 7 >> gpurun_out/${tag}_group.txt 2>&1
( cd /tmp && /root/repo/raytracer-795_b200/bin/raytracer $XML --gpus 8 ) > gpurun_out/${tag}_cli_gpus8.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --config c5 --steps 3 --warmup 3 > gpurun_out/${tag}_c5_n8.json 2> gpurun_out/${tag}_c5_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --config c2 --steps 10 --warmup 3 > gpurun_out/${tag}_c2_n8.json 2> gpurun_out/${tag}_c2_n8.err
