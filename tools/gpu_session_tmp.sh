for c in 1 0 1 0; do echo -n "CARVEOUT=$c: "; GGML_B200_CARVEOUT=$c python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sdxl 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['unet_eval_ms_batch16'])"; done
