source tools/gpu_runs/ab_variants.sh r2d_ab.log true
run nb nb --permille 10
run nb_pf16 nb_pf16 --permille 10
run nb_pf16_la6 nb_pf16_la6 --permille 10
run nb_s5 nb_s5 --permille 10
run nb_s5_d3 nb_s5_d3 --permille 10
run nb_pf8 nb_pf8 --permille 10
