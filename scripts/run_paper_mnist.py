"""dist_mnist_PAPER.yaml end to end (hetero split, 10-node cycle, DiNNO / DSGT / DSGD x 2000 rounds,
evaluation every 20 rounds) on the synthetic MNIST-shaped dataset; prints the accuracy-vs-round table
(the second half of BASELINE.json's metric: 'MNIST val-acc vs rounds')."""
import glob, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, yaml
from nn_distributed_training_b200.experiments import dist_mnist_ex
from nn_distributed_training_b200.visualization import load_results, rounds_to_threshold

with open(os.path.join(ROOT, "experiments", "dist_mnist_PAPER.yaml")) as f:
    conf = yaml.safe_load(f)
out = tempfile.mkdtemp()
conf["experiment"].update(output_metadir=out, data_dir="/nonexistent")
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
for pc in conf["problem_configs"].values():
    pc["optimizer_config"]["outer_iterations"] = rounds
p = os.path.join(out, "paper.yaml"); yaml.safe_dump(conf, open(p, "w"))
t0 = time.time(); dist_mnist_ex.experiment(p); torch.cuda.synchronize(); wall = time.time() - t0
run = glob.glob(os.path.join(out, "*_dist_mnist_PAPER"))[0]
res = load_results(run)
table = {}
for name, m in res.items():
    acc = torch.stack(m["top1_accuracy"]).mean(1)
    table[name] = {"final_mean_top1": float(acc[-1]), "top1_at_round": {str(k * 20): float(acc[k]) for k in (0, 5, 10, 25, 50, len(acc) - 2) if k < len(acc)},
                   "rounds_to_90pct": rounds_to_threshold(m, 0.90, 20), "rounds_to_97pct": rounds_to_threshold(m, 0.97, 20),
                   "final_consensus_max": float(m["consensus_error"][-1][1].max())}
print(json.dumps({"config": "dist_mnist_PAPER.yaml (synthetic MNIST, hetero split, 10-node cycle)", "rounds": rounds,
                  "wall_seconds_all_three_incl_eval_and_setup": wall, "results": table}))
