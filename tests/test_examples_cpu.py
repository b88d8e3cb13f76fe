"""Every example script runs to completion on CPU (small sizes) and prints what it promises -- the examples are the user-facing
counterparts of the reference's examples/ tree, so they must not rot."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    ("pretrain/train_gpt.py --steps 3 --layers 2 --hidden 64 --heads 4 --seq 64 --global-batch 4 --micro-batch 2", "final loss"),
    ("pretrain/train_hetu.py --config-name gpt_small_dp2_tp2 ds_parallel.dp=1 ds_parallel.tp=1 trainer.steps=2 trainer.bf16=false "
     "trainer.max_seq_length=32 model.n_positions=32 model.n_embd=32 model.n_layer=2 model.n_head=2 model.vocab_size=259 "
     "model.sequence_parallel=false ds_parallel.sequence_parallel=false", "steps 2"),
    ("sft/sft_lora.py", "adapter tensors"),
    ("sft/sft_hetu.py --config-name gpt_lora trainer.steps=6 sft.lora_rank=4 --prompt hello", "answer:"),
    ("malleus/replan.py", "estimated step time"),
    ("malleus/train_malleus.py --steps 5", "hetero path: False"),
    ("galvatron/search.py --gpus 8 --mem-gb 40", "galvatron_plan.json"),
    ("hydraulis/dynamic_dispatch.py", "makespan"),
    ("lobra/plan_and_dispatch.py --ngpus 16", "step estimate"),
    ("lobra/train_multi_lora.py", "mean loss"),
    ("ampelos/elastic_replan.py", "data loader resumes"),
    ("ctr/run_ctr.py --steps 3 --model deepfm --compress hash", "embedding compression hash"),
    ("gnn/run_gcn.py", "1.5-D layout"),
    ("hotspa/train_hot_switch.py", "strategy 1 seq 256"),
    ("moe/train_moe.py", "loss"),
    ("elastic/run_elastic.py", "generations"),
    ("efficiency/profile_attn.py --seq 256", "packed varlen"),
    ("bert/pretrain_bert.py --steps 11", "nsp-acc"),
    ("cnn/train_cnn.py --model resnet18 --steps 11 --batch 16 --width-div 8", "resnet18 step 10"),
    ("rec/train_ncf.py --steps 51", "auc"),
    ("nlp/train_transformer.py --steps 21 --batch 16 --seq 5 --vocab 12", "decoded"),
    ("hetero/convert_checkpoint.py examine .", "INCOMPLETE"),
    ("v1/train_ps_roles.py --steps 12 --port 0", "cosine(w, w*)"),
]


@pytest.mark.parametrize("cmd,expect", CASES, ids=[c[0].split()[0] + ("" if i < 3 else f"-{i}") for i, c in enumerate(CASES)])
def test_example_runs(cmd, expect, tmp_path):
    env = dict(os.environ, HETU_B200_FORCE_CPU="1", CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2", PYTHONPATH=ROOT, TRAINER_OUT=str(tmp_path / "out"))
    parts = cmd.split()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", parts[0])] + parts[1:], env=env, cwd=str(tmp_path), capture_output=True, text=True,
                       timeout=280)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert expect in out, out[-3000:]


def test_malleus_example_moves_a_running_job_onto_the_plan(tmp_path):
    """4 ranks: device 3 is reported 3x slower after step 4 -> batch shares 3 : 1, applied through the heterogeneous path"""
    env = dict(os.environ, HETU_B200_FORCE_CPU="1", CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1", PYTHONPATH=ROOT, TRAINER_OUT=str(tmp_path / "out"))
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nproc-per-node", "4", "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "examples", "malleus", "train_malleus.py"), "--slow-rank", "3", "--steps", "8"],
                       env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=400)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert '"micro-batches per pipeline": [3, 1]' in out and '"applied": "hetero"' in out and "hetero path: True" in out, out[-3000:]
