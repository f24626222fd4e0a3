L=mav_trajectory_generation_cmake_b200/lib
cp $L/libminsnap_b200.so $L/new.so
for v in base new base new; do cp $L/$v.so $L/libminsnap_b200.so; echo "== $v"; python tools/bench_long_chain.py 2>&1 | grep "chunked"; done
cp $L/new.so $L/libminsnap_b200.so
python tools/quick_bench.py --K 8 --B 524288 --steps 20
python tools/quick_bench.py --K 8 --B 131072 --steps 50
