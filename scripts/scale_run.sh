# kernel times per particle vs lattice size and vs how far the synthetic run has evolved (occupancy drifts from exactly one
# particle per cell towards Poisson-like): usage bash scripts/scale_run.sh "256 512" "2:3 6:3"   (warmup:steps pairs)
for n in ${1:-256 512}; do for ws in ${2:-2:3}; do
  w=${ws%%:*}; st=${ws##*:}
  python bench.py --ngrid $n --steps $st --warmup $w --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['kernels']; n=$n**3
print('N=$n warmup=$w steps=$st', ' '.join('%s %.3f ms %.1f ps/p' % (nm[:14], k[nm]['ms_per_step'], 1e9*k[nm]['ms_per_step']/n) for nm in ('projection_T00_Tij_project','kick_drift','rebin_sort')), 'total %.2f' % d['ms_per_step'])
"
done; done
