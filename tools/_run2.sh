for i in 1 2; do
for l in build_exp/lib_r1base.so noblegas_rtd_mcmc_b200/libngrtd.so build_exp/lib_swz_dadd.so; do
NGRTD_LIB=$PWD/$l python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$l', d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['sync_call']['value'], 'smp', d['sampler']['value'], d['ess']['ess_per_sec'])"
done; done
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv
