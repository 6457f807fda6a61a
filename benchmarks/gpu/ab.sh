# A/B of library variants on ONE box: serial kernel times (ktrace, 1 lane) and the pipelined step (3 lanes), twice each
export CUDA_DEVICE_MAX_CONNECTIONS=32
V=denovo-variant-caller-java_b200/lib/variants
for rep in 1 2; do
for name in "$@"; do
  export DNV_B200_LIB=$PWD/$V/$name.so
  s=$(DNV_LANES=1 python benchmarks/ktrace.py 6 | awk '/^ *[0-9]+ +0 /{n++; if(n>2){scr+=$4-$3; lst+=$6-$5; fin+=$8-$7; gap+=$7-$6; k++}} END{printf "serial: screen %.1f list %.1f gap %.1f finish %.1f", scr/k, lst/k, gap/k, fin/k}')
  p=$(python benchmarks/ktrace.py 24 | awk '/^ *[0-9]+ +[0-9] /{ lst+=$6-$5; fin+=$8-$7; k++} /steady step/{st=$3} END{printf "pipelined: step %s list %.1f finish %.1f", st, lst/k, fin/k}')
  echo "$name  $s | $p"
done; done
