# usage: ncu_any.sh TAG KERNEL_REGEX SKIP -- command...
mkdir -p gpurun_out
TAG=$1; REGEX=$2; SKIP=$3; shift 4
"$@" > gpurun_out/${TAG}_plain.json 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:$REGEX -s $SKIP -c 1 -f -o gpurun_out/${TAG} "$@" > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
