#!/bin/bash
# Round-2 record (second half): every GPU test with -s (measured parity values on disk), conv timings with and without
# epilogue statistics, the default bench line (rooflines, CUDA-eager and CPU baselines), the reference arm, the bench
# lines of configs 3-5, launch lists (CIFAR), ncu --set full of the dominant conv kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ATTNDM_PARITY_OUT=gpurun_out/parity_r02 timeout 1800 python -m pytest tests -q -s -m gpu -p no:cacheprovider > gpurun_out/r02_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r02_tests.log
timeout 300 python tools/conv_bench.py --shapes c128_32,c256_32,c128_16,c128_8,n256_32,out_32,in_32,c128_64,c128_128,c128_256 > gpurun_out/r02_conv_bench.log 2>&1
timeout 300 python tools/conv_bench.py --shapes c128_32,c256_32,c128_16,c128_8,c128_64 --stats 1 >> gpurun_out/r02_conv_bench.log 2>&1
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
echo "ref rc=$?"
for c in cifar10_w4_attn celeba_w8a8 church_w8a8; do
  timeout 900 python bench.py --config $c --steps 2 --warmup 3 --no-extras > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err
  echo "bench $c rc=$?"
done
bash tools/gpu_round2_s.sh > gpurun_out/r02_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qconv_i8_halo -s 3 -c 1 \
    -f -o gpurun_out/prof_conv_r02b python tools/conv_bench.py --graph 0 --shapes c128_32 --iters 3 > gpurun_out/r02_ncu_conv.log 2>&1
echo "ncu conv rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" 2>&1 | tail -2
tail -3 gpurun_out/r02_tests.log; cat gpurun_out/r02_conv_bench.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'keep_last', d['e2e_keep_last']['value'], 'roof', d['roofline']['frac'], d['roofline']['ms_per_launch'], d['roofline'].get('ms_per_launch_eager'))
for r in d.get('roofline_hbm', []): print(r['kernel'][:50], round(r['achieved']), round(r['frac'],3))
print(d['cuda_eager_baseline']); print(d['cpu_baseline']); print(d['clocks'])
print(open('gpurun_out/r02_bench_ref.json').read()[:400])
for c in ['cifar10_w4_attn','celeba_w8a8','church_w8a8']:
    try:
        e=json.loads(open(f'gpurun_out/r02_bench_{c}.json').read().strip().splitlines()[-1])
        print(c, round(e['value'],1), 'img/s', round(e['ms_per_step'],1), 'ms/pass', e['launches_per_denoising_step'], 'launches', 'e2e', round(e['e2e']['value'],1))
    except Exception as ex:
        print(c, 'failed', ex)
PY
