# usage: bash scripts/abl_run.sh "variant names ('-' = product build)" "sizes" "warmup:steps ..."
for v in ${1:--}; do
  if [ "$v" = "-" ]; then unset GEVB_LIB; else export GEVB_LIB=$PWD/build/abl/libgevb_$v.so; fi
  echo "== variant $v"; bash scripts/scale_run.sh "${2:-512}" "${3:-2:3 8:3}"
done
