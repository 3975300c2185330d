#!/bin/sh
# final round-1 captures of the regime-sorting kernel and the density kernel
sh tools/ncu_one.sh final_binomial_d16 --dist binomial --streams 1048576 --draws 16
sh tools/ncu_one.sh final_poisson_d1 --dist poisson --streams 4194304 --draws 1
sh tools/ncu_one.sh final_hypergeometric_d16 --dist hypergeometric --streams 524288 --draws 16
sh tools/ncu_one.sh final_density_poisson --workload density --density poisson --obs 50000000
