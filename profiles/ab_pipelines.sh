# throughput against the number of alignments in flight (run on the GPU box): "pipelines host-threads" pairs
[ $# -eq 0 ] && set -- "16 2" "32 2" "32 4" "64 4"
for cfg in "$@"; do
set -- $cfg
python bench.py --no-cpu-baseline --no-extra --windows 2 --pipelines $1 --host-threads $2 --batch 64 > gpurun_out/ab_p$1_t$2.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/ab_p$1_t$2.json')); print('pipelines $1 threads $2', round(d['value']), round(d['e2e']['value']))"
done
