run() { echo "CTAS=$1 SA=$2 T=$3 :: $(OS3D_SPCONV_CTAS=$1 OS3D_SPCONV_SA=$2 OS3D_SPCONV_TILES=$3 timeout 60 python tools/run_spconv.py $4 | cut -d: -f2 | cut -d, -f1)  [$4]"; }
run 2 8 1 "3 192 192"; run 1 8 2 "3 192 192"; run 1 4 2 "3 192 192"; run 1 8 1 "3 192 192"
run 2 8 1 "3 384 192"; run 1 8 2 "3 384 192"; run 1 4 2 "3 384 192"
run 2 8 2 "2 96 96"; run 1 8 4 "2 96 96"; run 1 8 5 "2 96 96"
run 2 8 2 "2 192 96"; run 1 8 4 "2 192 96"
run 2 8 5 "1 48 48"; run 1 8 5 "1 48 48"; run 1 8 10 "1 48 48"
