# A/B of library builds through bench.py (run on the GPU box): SM_B200_LIB selects the .so
for lib in "$@"; do
SM_B200_LIB=$PWD/staticmapping_b200/$lib python bench.py --no-cpu-baseline --no-extra --windows 2 > gpurun_out/ab_$lib.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/ab_$lib.json')); print('$lib', round(d['value']), round(d['e2e']['value']), round(d['latency']['ms_per_alignment_device'],3), {k: round(x,3) for k,x in d['roofline']['per_alignment_ms'].items()})"
done
