# usage: ab.sh REPS "ENV1" "ENV2" ...  -- interleaved repetitions of bench.py under each environment
REPS=$1; shift
B="python bench.py --steps 30 --warmup 5 --skip-cpu --skip-e2e --skip-aligned --skip-long"
i=0
for r in $(seq 1 $REPS); do
  k=0
  for e in "$@"; do
    env $e $B > gpurun_out/ab_${k}_r${r}.log 2>&1
    k=$((k+1))
  done
done
echo done
