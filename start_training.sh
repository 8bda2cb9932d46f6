#!/bin/bash
# Launcher: KEY=VALUE interface.
#   sh start_training.sh MASTER_ADDR=127.0.0.1 MASTER_PORT=8888 N_NODES=1 GPUS_PER_NODE=8 NODE_RANK=0 \
#       WORKSPACE=/path/ws DATASET=llff VERSION=exp1 EXTRA_CONFIG='{"training.gpus": "0,1,2,3,4,5,6,7"}'
# DATASET in {llff, flowers, kitti_raw, dtu, realestate10k(default)}.
set -e
cd "$(dirname "$0")"

declare -A OPT=( [MASTER_ADDR]=127.0.0.1 [MASTER_PORT]=8888 [N_NODES]=1 [GPUS_PER_NODE]=1 [NODE_RANK]=0
                 [WORKSPACE]="" [DATASET]=realestate10k [VERSION]="" [EXTRA_CONFIG]="{}" )
for kv in "$@"; do
    key="${kv%%=*}"; val="${kv#*=}"
    if [[ -n "${OPT[$key]+x}" ]]; then OPT[$key]="$val"; else echo "ignoring unknown argument: $key" >&2; fi
done
for k in MASTER_ADDR MASTER_PORT N_NODES GPUS_PER_NODE NODE_RANK WORKSPACE DATASET VERSION EXTRA_CONFIG; do
    echo "$k: ${OPT[$k]}"
done
[[ -z "${OPT[WORKSPACE]}" || -z "${OPT[VERSION]}" ]] && { echo "WORKSPACE and VERSION are required" >&2; exit 2; }

case "${OPT[DATASET]}" in
    llff|flowers|kitti_raw|dtu) PARAMS="./configs/params_${OPT[DATASET]}.yaml" ;;
    *)                          PARAMS="./configs/params_realestate.yaml" ;;
esac
echo "default params: $PARAMS"

exec python3 -m torch.distributed.run \
    --master-addr "${OPT[MASTER_ADDR]}" --master-port "${OPT[MASTER_PORT]}" \
    --nnodes "${OPT[N_NODES]}" --nproc-per-node "${OPT[GPUS_PER_NODE]}" --node-rank "${OPT[NODE_RANK]}" \
    train.py --config_path "$PARAMS" --workspace "${OPT[WORKSPACE]}" --version "${OPT[VERSION]}" \
    --extra_config "${OPT[EXTRA_CONFIG]}"
