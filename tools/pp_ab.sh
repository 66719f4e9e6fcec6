for v in "$@"; do echo "== $v"; DSC_LIB_PATH=deepsc-gan_b200/csrc/_var/lib_$v.so timeout 100 python tools/pp_check.py ${MODE:-time-only} 2>&1 | grep -E "one-tile|identical|MISMATCH" | head -8; done
