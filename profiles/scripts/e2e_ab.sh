# A/B of the end-to-end leg: layer image on / off (default chunk schedule)
for i in 1 2 3; do
  for img in 1 0; do
    S5B200_LAYER_IMAGE=$img timeout 100 python bench.py --steps 10 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('image $img', round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,2), round(d['e2e']['stream_ordered_value']/1e6,2))"
  done
done
