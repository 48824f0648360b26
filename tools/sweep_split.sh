for T in 2 3 4 6; do for B in "1024 10" "128 2" "1024 30" "1024 100"; do set -- $B; MINA_B200_SPLIT_TARGET=$T python bench.py --steps 6 --warmup 3 --batch $1 --corrupt $2 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.readline());print('T=$T batch=$1 corrupt=$2',round(d['ms_per_step'],2),'ms',d['roofline']['msms_per_step'],'msms')"; done; done
