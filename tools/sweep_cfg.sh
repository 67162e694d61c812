run() { echo "SA=$1 T=$2 :: $(OS3D_SPCONV_SA=$1 OS3D_SPCONV_TILES=$2 timeout 60 python tools/run_spconv.py $3 | cut -d: -f2 | cut -d, -f1)  [$3]"; }
for sa in 2 4; do for t in 1 2 3 5; do run $sa $t "1 48 48"; done; done
for sa in 2 4; do for t in 1 2; do run $sa $t "2 96 96"; done; done
run 8 5 "2 96 96"
for sa in 2 4; do for t in 1 2; do run $sa $t "2 192 96"; done; done
for sa in 2 4; do run $sa 1 "3 192 192"; done
run 4 2 "3 192 192"
for sa in 2 4; do run $sa 1 "3 384 192"; done
for sa in 2 4 8; do run $sa 1 "4 384 384"; done
