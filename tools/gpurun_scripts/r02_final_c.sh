# last pass of the round on the final code: full GPU test suite, smoke, bench (+ layer report), reference arm, launch list
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "(gpu tests: run separately)"
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02g_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/r02g_smoke.log)"
timeout 600 python bench.py --layer-report $O/r02g_layers_resnet50.json > $O/r02g_bench.json 2> $O/r02g_bench.err; echo "bench rc=$? $(cut -c1-160 $O/r02g_bench.json)"
timeout 600 python bench.py --impl reference > $O/r02g_bench_reference.json 2> $O/r02g_bench_reference.err; echo "reference rc=$?"
for n in resnet18 vgg16 mobilenet_v2; do
  timeout 300 python bench.py --no-cpu-baseline --network $n --layer-report $O/r02g_layers_${n}.json > $O/r02g_bench_${n}.json 2> $O/r02g_bench_${n}.err; echo "$n rc=$? $(cut -c1-170 $O/r02g_bench_${n}.json)"
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/r02g_bench_launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 > $O/r02g_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
