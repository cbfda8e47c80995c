# alternate two library builds (VYOLO_LIB_VARIANT) over the same timing script: bash tools/ab_variants.sh old "" 3
a=$1; b=$2; n=${3:-3}
for i in $(seq $n); do
  for v in "$a" "$b"; do
    echo "== variant '${v}' run $i"; VYOLO_LIB_VARIANT=$v timeout 120 python tools/str_nohits.py 2>&1 | cut -c1-75
  done
done
