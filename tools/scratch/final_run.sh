python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/prof_round.sh r02g
bash tools/bench_configs.sh r02g
