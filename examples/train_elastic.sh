#!/bin/bash
# BASELINE config 5: four workers launched independently (2 GPUs each). Kill any worker's torchrun: the others drop it at
# the next outer boundary after the heartbeat timeout. Run `./train_elastic.sh worker 2` again: it is re-admitted and receives
# master/Adam/θ₀/momentum from a survivor by broadcast. `./train_elastic.sh status` prints the membership.
set -euo pipefail
cd "$(dirname "$0")/.."
export GLOBAL_ADDR=127.0.0.1 GLOBAL_PORT="${GLOBAL_PORT:-29400}"
case "${1:-all}" in
  serve)  exec python -m prime_b200.parallel.elastic serve --port "$GLOBAL_PORT" ;;
  status) exec python -m prime_b200.parallel.elastic status --port "$GLOBAL_PORT" ;;
  worker)
    w="$2"
    GLOBAL_UNIQUE_ID="w$w" CUDA_VISIBLE_DEVICES="$((2*w)),$((2*w+1))" exec python -m torch.distributed.run --nnodes=1 \
        --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$((29510+w))" -m diloco.train @configs/1B/elastic.toml "${@:3}" ;;
  all)
    python -m prime_b200.parallel.elastic serve --port "$GLOBAL_PORT" & store=$!
    trap 'kill $store $(jobs -p) 2>/dev/null' EXIT
    sleep 2
    for w in 0 1 2 3; do "$0" worker "$w" & done
    wait ;;
esac
