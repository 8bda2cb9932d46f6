#!/bin/sh
# Launcher: KEY=VALUE interface (POSIX sh, so `sh start_training.sh ...` works with dash as well as bash).
#   sh start_training.sh MASTER_ADDR=127.0.0.1 MASTER_PORT=8888 N_NODES=1 GPUS_PER_NODE=8 NODE_RANK=0 \
#       WORKSPACE=/path/ws DATASET=llff VERSION=exp1 EXTRA_CONFIG='{"training.gpus": "0,1,2,3,4,5,6,7"}'
# DATASET in {llff, flowers, kitti_raw, dtu, realestate10k(default)}.
set -e
cd "$(dirname "$0")"

MASTER_ADDR=127.0.0.1; MASTER_PORT=8888; N_NODES=1; GPUS_PER_NODE=1; NODE_RANK=0
WORKSPACE=""; DATASET=realestate10k; VERSION=""; EXTRA_CONFIG="{}"
for kv in "$@"; do
    key="${kv%%=*}"; val="${kv#*=}"
    case "$key" in
        MASTER_ADDR)   MASTER_ADDR="$val" ;;
        MASTER_PORT)   MASTER_PORT="$val" ;;
        N_NODES)       N_NODES="$val" ;;
        GPUS_PER_NODE) GPUS_PER_NODE="$val" ;;
        NODE_RANK)     NODE_RANK="$val" ;;
        WORKSPACE)     WORKSPACE="$val" ;;
        DATASET)       DATASET="$val" ;;
        VERSION)       VERSION="$val" ;;
        EXTRA_CONFIG)  EXTRA_CONFIG="$val" ;;
        *) echo "ignoring unknown argument: $key" >&2 ;;
    esac
done
echo "MASTER_ADDR: $MASTER_ADDR"; echo "MASTER_PORT: $MASTER_PORT"; echo "N_NODES: $N_NODES"
echo "GPUS_PER_NODE: $GPUS_PER_NODE"; echo "NODE_RANK: $NODE_RANK"; echo "WORKSPACE: $WORKSPACE"
echo "DATASET: $DATASET"; echo "VERSION: $VERSION"; echo "EXTRA_CONFIG: $EXTRA_CONFIG"
if [ -z "$WORKSPACE" ] || [ -z "$VERSION" ]; then echo "WORKSPACE and VERSION are required" >&2; exit 2; fi

case "$DATASET" in
    llff|flowers|kitti_raw|dtu) PARAMS="./configs/params_${DATASET}.yaml" ;;
    *)                          PARAMS="./configs/params_realestate.yaml" ;;
esac
echo "default params: $PARAMS"

exec python3 -m torch.distributed.run \
    --master-addr "$MASTER_ADDR" --master-port "$MASTER_PORT" \
    --nnodes "$N_NODES" --nproc-per-node "$GPUS_PER_NODE" --node-rank "$NODE_RANK" \
    train.py --config_path "$PARAMS" --workspace "$WORKSPACE" --version "$VERSION" \
    --extra_config "$EXTRA_CONFIG"
