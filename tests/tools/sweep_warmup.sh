mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python tests/probe_parity.py gpurun_out/pp_$tag.json 21 metop_ahrpt psk8 jpss_hrd bpsk_half > gpurun_out/pp_$tag.log 2>&1; echo "== $tag $@"; grep -v "^{" gpurun_out/pp_$tag.log | cut -c1-420; }
run mmw125 B200_MM_TOL=0.004 B200_MM_WARMUP_SCALE=1.25 B200_COSTAS_TOL=2.5e-6
run mmw150 B200_MM_TOL=0.004 B200_MM_WARMUP_SCALE=1.5 B200_COSTAS_TOL=2.5e-6 B200_COSTAS_WARMUP_SCALE=0.75
run mmw200 B200_MM_TOL=0.002 B200_MM_WARMUP_SCALE=2.0 B200_COSTAS_TOL=2.5e-6 B200_COSTAS_WARMUP_SCALE=0.5
