# GPU experiment: the ordered mode's end-to-end number (bench.py e2e) against the rows per launch of the host
# call and the SMs k_own leaves to the plan of the next chunk.   usage: bash tools/e2e_sweep.sh
for args in "--e2e-chunk-rows 4194304" "--e2e-chunk-rows 6291456" "--e2e-chunk-rows 8388608" "--e2e-chunk-rows 4194304 --opt own_spare_sms=4" "--e2e-chunk-rows 4194304 --opt own_spare_sms=12" "--e2e-chunk-rows 6291456 --opt own_spare_sms=12"; do
timeout 300 python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline --no-other-configs --seam-rows 0 --parity-rows 0 $args 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$args', 'value', round(d['value']/1e6), 'e2e', round(d['e2e']['value']/1e6), 'compact', round(d['e2e']['compact_h2d']['value']/1e6))"
done
