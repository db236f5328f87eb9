source tools/gpu_runs/ab_variants.sh r2c_ab.log true
run s2w4 s2w4 --permille 10 --warps 4
run s3w5 s3w5 --permille 10 --warps 5 --reads 71040
run s3w8 s3w8 --permille 10
run d1 d1 --permille 10
run pf16 pf16 --permille 10
run la6 la6 --permille 10
