run() { name=$1; shift; timeout 100 python bench.py --steps 8 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; python -c "
import json;d=json.load(open('gpurun_out/ab_$name.json'));print('$name',round(d['value']),round(d['ms_per_step'],2),round(d['e2e']['value']),d['agreement']);print(d['ms_per_step_by_kernel'])"; }
