# end-to-end leg of bench.py under different clip-chunk schedules of s5b200_model_forward_host (S5B200_E2E_CHUNKS)
for sched in "" "" "1,2,4,6,6,6,4,2,1" "1,2,4,6,6,6,4,2,1" "4,4,4,4,4,4,4,4" "8,8,8,8" "16,16"; do
  S5B200_E2E_CHUNKS=$sched timeout 100 python bench.py --steps 10 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('sched [$sched]', round(d['e2e']['value']/1e6,2), round(d['e2e']['stream_ordered_value']/1e6,2))"
done
S5B200_LAYER_IMAGE=0 timeout 100 python bench.py --steps 10 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('no layer image', round(d['e2e']['value']/1e6,2), round(d['e2e']['stream_ordered_value']/1e6,2))"
