"""Elastic training controller demo: 4 workers, one dies in the first generation, the controller re-plans to the
survivors and restarts them (workers would resume from the latest ModelSaver checkpoint)."""
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from hetu_b200.rpc import ElasticServer, ElasticStrategy

worker = tempfile.NamedTemporaryFile("w", suffix=".py", delete=False)
worker.write("import sys, time\ngen, rank, world = map(int, sys.argv[1:4])\nprint(f'gen {gen} rank {rank}/{world} running', flush=True)\n"
             "if gen == 0 and rank == 3: sys.exit(1)\ntime.sleep(0.5)\n")
worker.close()


def launch(gen, plan, addr):
    return [subprocess.Popen([sys.executable, worker.name, str(gen), str(r), str(plan["num_devices"])]) for r in range(plan["num_devices"])]


es = ElasticServer(launch, 4, ElasticStrategy(tp=2), port=24700, max_restarts=2)
print("exit", es.run(), "generations", [(g["gen"], g["plan"]) for g in es.generations])
