source tools/gpu_runs/ab_variants.sh r2b_ab.log true
run base base --permille 10,5,7,15
run la2 la2 --permille 10
run pf64 pf64 --permille 10
run w16 w16 --permille 10 --warps 16
run w20 w20 --permille 10 --warps 20 --reads 94720
run prof prof --permille 10
