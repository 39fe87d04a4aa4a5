mkdir -p gpurun_out/v9
for blk in 512 592 1024 1184; do for geom in 1 2; do
TMB_PIPE_GEOM=$geom python bench.py --steps 6 --no-cpu --block $blk 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('block $blk geom $geom: value %.0f e2e %.0f ms/step %.3f tfce %.3f fit %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['fit']['ms_per_launch']))"
done; done
