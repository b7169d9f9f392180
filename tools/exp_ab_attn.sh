for i in 1 2; do
for lib in libctta.so libctta_attnpk.so; do
  CTTA_LIB=$PWD/consistencytta_b200/$lib python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-100 | sed "s/^/$lib /"
done; done
