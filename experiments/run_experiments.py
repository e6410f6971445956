"""Monte-Carlo experiment driver on the B200 path -- the recipe of the reference's experiments/run_experiments.py
(:38-47,102-215): N_MCMC_RUNS repeats x 3 L-kernel configurations, `10*(i+1)` seeds, CSV outputs with the same file
names and layout (np.savetxt of mean_estimate, variance_estimate, ess, phi, acceptance_rate), plus the MSE-vs-truth
summary of experiments/plot_experiments.py:61-79 (no plotting).

    python experiments/run_experiments.py --model arma --runs 25 --N 100 --K 15 --out experiments/output
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "smc-nuts_b200"))

CONFIGS = [("forward_lkernel", "forwardsLKernel", False), ("gaussian_lkernel", "GaussianApproxLKernel", False),
           ("asymptotic_lkernel", "asymptoticLKernel", True)]


def load_truth(model_dir, name):
    """Column 2 of `<model>.params` = gold-standard posterior means (run_experiments.py:68-76)."""
    return np.array([float(line.split()[1]) for line in (model_dir / f"{name}.params").read_text().splitlines() if line.strip()])


def save_output(smc, strategy, i, output_dir):
    """run_experiments.py:195-215."""
    path = Path(output_dir) / strategy
    path.mkdir(parents=True, exist_ok=True)
    np.savetxt(path / f"mean_estimate_{i}.csv", smc.mean_estimate, delimiter=",")
    np.savetxt(path / f"var_estimate_{i}.csv", smc.variance_estimate, delimiter=",")
    np.savetxt(path / f"ess_{i}.csv", smc.ess, delimiter=",")
    np.savetxt(path / f"phi_{i}.csv", smc.phi, delimiter=",")
    np.savetxt(path / f"acceptance_rate_{i}.csv", smc.acceptance_rate, delimiter=",")


def mse_mean_var(x, ground_truth):
    """plot_experiments.py:60-78: x is [M runs, K+1 iterations, D]; the squared error of every run and iteration is
    averaged over the D parameters, then its mean and its (population, ddof = 0) variance are taken over the M runs.
    Returns (mse_mean[K+1], mse_var[K+1])."""
    x = np.asarray(x, dtype=np.float64)
    mse_per_run_iter = np.mean(np.square(x - np.asarray(ground_truth, dtype=np.float64)[None, None, :]), axis=2)
    return np.mean(mse_per_run_iter, axis=0), np.var(mse_per_run_iter, axis=0)


def mse_per_iteration(means, truth):
    """First output of mse_mean_var: per-iteration squared error averaged over runs and parameters."""
    return mse_mean_var(means, truth)[0]


def run(model_name="arma", runs=25, N=100, K=15, out=None, configs=CONFIGS, verbose=True):
    from smcnuts.distributions import StdNormal
    from smcnuts.model.bridgestan import StanModel
    from smcnuts.model.device_model import DATA_DIR
    from smcnuts.smc_sampler import SMCSampler
    model_dir = DATA_DIR / model_name
    step_size = json.loads((model_dir / "model_config.json").read_text()).get("step_size", 0.5)   # run_experiments.py:79-90
    target = StanModel(model_name, str(model_dir / f"{model_name}.stan"), str(model_dir / f"{model_name}.json"))
    truth = load_truth(model_dir, model_name)
    results = {}
    for strategy, lkernel, tempering in configs:
        means = []
        for i in range(runs):
            seed = 10 * (i + 1)                                                                    # run_experiments.py:106
            smc = SMCSampler(K=K, N=N, target=target, step_size=step_size, sample_proposal=StdNormal(target.dim),
                             momentum_proposal=StdNormal(target.dim), lkernel=lkernel, tempering=tempering, rng=seed)
            smc.sample(show_progress=False)
            if out:
                save_output(smc, strategy, i, Path(out) / model_name)
            means.append(smc.mean_estimate)
        mse, mse_var = mse_mean_var(means, truth)
        results[strategy] = dict(means=np.array(means), mse=mse, mse_var=mse_var)
        if verbose:
            print(f"{model_name} {strategy}: final-iteration MSE vs gold means = {results[strategy]['mse'][-1]:.3e} "
                  f"({runs} runs, N={N}, K={K})")
    return results, truth


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="arma", choices=["arma", "PRMwCD"])
    ap.add_argument("--runs", type=int, default=25)
    ap.add_argument("--N", type=int, default=100)
    ap.add_argument("--K", type=int, default=15)
    ap.add_argument("--out", default=str(ROOT / "experiments" / "output"))
    a = ap.parse_args()
    run(a.model, a.runs, a.N, a.K, a.out)
