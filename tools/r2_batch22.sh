#!/bin/bash
# per-instruction ncu captures of the three NUTS kernels with the final build of the round
O=gpurun_out/r2w; mkdir -p $O
for w in "gauss 18" "PRMwCD 17" "arma 20"; do
  n=$(echo $w | cut -d' ' -f1)
  timeout 300 python tools/ab_time.py $w 1 > $O/ab_${n}_plain.log 2>&1 && \
    timeout 600 ncu --set full --import-source on --clock-control none -k regex:nuts_transition -s 2 -c 1 -o $O/${n}_prof \
        python tools/ab_time.py $w 1 > $O/ncu_$n.log 2>&1
  [ -f $O/${n}_prof.ncu-rep ] && ncu -i $O/${n}_prof.ncu-rep --page details --csv > $O/${n}_details.csv 2>/dev/null
  [ -f $O/${n}_prof.ncu-rep ] && ncu -i $O/${n}_prof.ncu-rep --page source --csv > $O/${n}_src.csv 2>/dev/null
  [ -f $O/${n}_prof.ncu-rep ] && ncu -i $O/${n}_prof.ncu-rep --page raw --csv > $O/${n}_raw.csv 2>/dev/null
  rm -f $O/${n}_prof.ncu-rep
done
cat $O/ab_*_plain.log; ls -la $O
