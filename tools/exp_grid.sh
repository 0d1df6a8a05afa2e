for v in "4 8" "4 16" "2 8" "2 16" "1 8"; do set -- $v
  echo "== step $1 pop $2: $(FCS_PHMM_COARSE_STEP=$1 FCS_PHMM_POP_DIV=$2 python tools/quick_bench.py --cfg c3 --e2e --iters 7 2>&1 | grep -i 'e2e' | cut -c1-60)"
done
