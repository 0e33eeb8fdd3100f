# Round-2 evidence run (one B200): headline bench, reference arm, ncu launch list, ncu --set full of the two streaming kernels
set -x
python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-latency > gpurun_out/r2_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:out_heads -s 4 -c 1 -o gpurun_out/r2_prof_out_heads \
    python tools/time_sections.py 18944 2000 vae - u16 > gpurun_out/r2_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:enc_first_fwd -s 4 -c 1 -o gpurun_out/r2_prof_enc_fwd \
    python tools/time_sections.py 18944 2000 vae - u16 > gpurun_out/r2_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:latent_block_bwd -s 4 -c 1 -o gpurun_out/r2_prof_latent_bwd \
    python tools/time_sections.py > gpurun_out/r2_ncu_c.log 2>&1
tail -c 600 gpurun_out/r2_bench_1gpu.err
