set -x
for defs in "" "-DSISUA_PM_POLY_LINKS=0" "-DSISUA_PM_POLY_LINKS=0 -DSISUA_TC_NO_BACKOFF" "-DSISUA_PM_POLY_LINKS=1 -DSISUA_PM_POLY_PI=1"; do
  SISUA_NVCC_DEFS="$defs" python -m sisua_b200.build --force > /dev/null 2>&1
  SISUA_NVCC_DEFS="$defs" python tools/time_sections.py >> gpurun_out/r2_variants.txt 2>&1
done
SISUA_NVCC_DEFS="-DSISUA_PM_POLY_LINKS=0" python -m sisua_b200.build --force > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:out_heads -s 3 -c 1 -o gpurun_out/prof_out_heads_r2a python tools/time_sections.py > gpurun_out/ncu_r2a.log 2>&1
cat gpurun_out/r2_variants.txt
