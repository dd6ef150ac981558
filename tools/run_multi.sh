#!/bin/bash
# One box, N GPUs: the scaling runs of a round (bench step + e2e legs, configs[2] corpus pass in both modes, configs[4]
# sweep).  Usage: tools/run_multi.sh N [tag]   -> gpurun_out/{bench,corpus_dev,corpus_host,sweep}_<tag>N.json
N=${1:-8}; TAG=${2:-r2_}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
P=29500
run() { name=$1; shift; "$@" > gpurun_out/${TAG}${name}_${N}gpu.json 2> gpurun_out/${TAG}${name}_${N}gpu.err; echo "$name rc=$?"; tail -c 1500 gpurun_out/${TAG}${name}_${N}gpu.json; echo; }
if [ "$N" = 1 ]; then TR="python"; PORT() { true; }; fi
run bench $TR $([ "$N" != 1 ] && echo --master-port $((P+1))) bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --sustained-seconds 2
run corpus_dev $TR $([ "$N" != 1 ] && echo --master-port $((P+2))) tools/run_corpus.py --mode device
run corpus_host $TR $([ "$N" != 1 ] && echo --master-port $((P+3))) tools/run_corpus.py --mode host
run sweep $TR $([ "$N" != 1 ] && echo --master-port $((P+4))) tools/run_configs.py --sweep-only --out gpurun_out/${TAG}sweep_rows_${N}gpu.json
nvidia-smi topo -m > gpurun_out/${TAG}topo_${N}gpu.txt 2>&1; numactl -H >> gpurun_out/${TAG}topo_${N}gpu.txt 2>&1; lscpu | head -20 >> gpurun_out/${TAG}topo_${N}gpu.txt
