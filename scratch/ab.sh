for rep in 1 2; do
for d in . scratch/wt; do
  echo "== $d"
  (cd $d && B200FA_NO_REBUILD=1 python profiles/microbench.py --hq 32 --hk 32 --nkv 4096 2>&1 | tail -1 | cut -c1-70
   B200FA_NO_REBUILD=1 python profiles/microbench.py --hq 32 --hk 32 --nkv 2048 --nq 2048 --causal --qf16 2>&1 | tail -1 | cut -c1-90
   B200FA_NO_REBUILD=1 python profiles/microbench.py --hq 32 --hk 32 --nkv 8192 --nq 8192 --causal --qf16 --steps 50 2>&1 | tail -1 | cut -c1-90)
done
done
