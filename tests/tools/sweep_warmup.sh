# Diagnostic: junction residuals / parity statistics of the loop kernels against warm-up length, gear-shift fraction and tolerances
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python tests/probe_parity.py gpurun_out/pp_$tag.json 21 metop_ahrpt bpsk_half jpss_hrd psk8 > gpurun_out/pp_$tag.log 2>&1; echo "== $tag $@"; grep -v "^{" gpurun_out/pp_$tag.log | cut -c1-330; }
run g0 B200_GEAR_SCALE=0
run g25 B200_GEAR_SCALE=0.25
run g40 B200_GEAR_SCALE=0.4
run g25w80 B200_GEAR_SCALE=0.25 B200_MM_WARMUP_SCALE=0.8 B200_COSTAS_WARMUP_SCALE=0.75
run g33w60 B200_GEAR_SCALE=0.33 B200_MM_WARMUP_SCALE=0.6 B200_COSTAS_WARMUP_SCALE=0.5
run g25t4 B200_GEAR_SCALE=0.25 B200_MM_TOL=0.004
