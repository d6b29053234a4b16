for v in "" regcol3 minb3 regcol4; do
  if [ -z "$v" ]; then unset FOCK_B200_LIB; else export FOCK_B200_LIB=$PWD/scratch/variants/libfock_$v.so; fi
  echo "== variant ${v:-default}"; python tools/tune_slos.py 12 24 2>&1 | tail -1
done
unset FOCK_B200_LIB
for v in "" ccminb5; do
  if [ -z "$v" ]; then unset FOCK_B200_LIB; else export FOCK_B200_LIB=$PWD/scratch/variants/libfock_$v.so; fi
  python tools/time_cc.py 65536 2>&1 | tail -1
done
unset FOCK_B200_LIB
for v in "" glynn4; do
  if [ -z "$v" ]; then unset FOCK_B200_LIB; else export FOCK_B200_LIB=$PWD/scratch/variants/libfock_$v.so; fi
  python tools/tune_perm.py 2>&1 | tail -1
done
unset FOCK_B200_LIB
