for st in 1 2 4; do
  echo "== CTTA_STREAM_ST=$st"
  for taps in 3 7 11; do for kind in c1 c2h; do
    CTTA_STREAM_ST=$st python tools/run_one_gemm.py conv1d --c 128 --taps $taps --dil 1 --rows 40968 --batch 64 --kind $kind --iters 20 --seconds 1.0
  done; done
done
