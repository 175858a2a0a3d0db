# Mirrors the reference's `make infer` (Makefile:14-21 there) on the B200 engine.
TEXT ?= "Hello World and goodbye"
SOURCE ?= "style.pt"
EXP ?= "data/best_exp"
CONFIG ?= ""
CHECKPOINT ?= ""
OUTPUT ?= "prediction"
PKG := diffusion-handwriting-generation.pytorch_b200

.PHONY: build infer test test-gpu bench

build:
	python -c "import __graft_entry__ as g; g.build()"

infer: build
	PYTHONPATH="$(PKG)" python -m dhg_b200.inference \
		--prompt=$(TEXT) \
		--source=$(SOURCE) \
		--experiment_path=$(EXP) \
		--config_path=$(CONFIG) \
		--checkpoint_path=$(CHECKPOINT) \
		--output=$(OUTPUT)

test:
	python -m pytest tests -x -q -m "not gpu"

test-gpu:
	python -m pytest tests -x -q -m gpu

bench:
	python bench.py
