# threads per model (P) of k_models on the C5 grid: P = 64 puts 32 consecutive segments of ONE light curve into a warp, P = 4 four
# adjacent segments of each of 8 models (all at the same phase of their light curves)
for P in 64 32 16 8 4; do echo "P=$P"; MLM_MODELS_P=$P python tools/profile_models.py 10000 3 2>&1 | tail -1; done
for P in 8 4; do for seg in 16 63; do echo "P=$P seg=$seg"; MLM_SEGMENT=$seg MLM_MODELS_P=$P python tools/profile_models.py 10000 3 2>&1 | tail -1; done; done
MLM_MODELS_P=4 python tools/count_events_models.py 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print(d['segment'], d['active_lanes'])"
