"""Generate tests/golden/*.npz by executing the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python -m oracle.make_golden            # regenerates every fixture

The reference has no tests or golden vectors of its own, so
the fixtures are outputs of the reference's own functions on seeded inputs.
Each fixture stores the inputs too, so the GPU-box tests need nothing else.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def gae_inputs(T, N, seed=0, p_term=0.01, p_tl=0.005):
    """SURVEY.md section 8(d) recipe, in this exact draw order."""
    rs = np.random.RandomState(seed)
    values = rs.randn(T, N, 1)
    rewards = rs.randn(T, N, 1)
    terminals = (rs.rand(T, N, 1) < p_term)
    time_limits = (rs.rand(T, N, 1) < p_tl)
    last_value = rs.randn(N, 1)
    return values, rewards, terminals, time_limits, last_value


def golden_gae(trl):
    from torchrl.replay_buffers import OnPolicyReplayBuffer
    out = {}
    cases = [(128, 8, 0, 0.01, 0.005), (37, 5, 1, 0.2, 0.1), (1, 3, 2, 0.5, 0.5), (256, 33, 3, 0.02, 0.02)]
    for ci, (T, N, seed, pt, ptl) in enumerate(cases):
        values, rewards, terminals, time_limits, last_value = gae_inputs(T, N, seed, pt, ptl)
        for flt in (True, False):
            buf = OnPolicyReplayBuffer(max_replay_buffer_size=T * N, env_nums=N, time_limit_filter=flt)
            buf._values = values.astype(np.float64)
            buf._rewards = rewards.astype(np.float64)
            buf._terminals = terminals.astype(np.float64)
            buf._time_limits = time_limits.astype(np.float64)
            buf.generalized_advantage_estimation(last_value, 0.99, 0.95)
            tag = "c%d_f%d" % (ci, int(flt))
            out[tag + "_gae_advs"] = buf._advs.copy()
            out[tag + "_gae_rets"] = buf._estimate_returns.copy()
            buf.discount_reward(last_value, 0.99)
            out[tag + "_disc_advs"] = buf._advs.copy()
            out[tag + "_disc_rets"] = buf._estimate_returns.copy()
        out["c%d_values" % ci] = values
        out["c%d_rewards" % ci] = rewards
        out["c%d_terminals" % ci] = terminals
        out["c%d_time_limits" % ci] = time_limits
        out["c%d_last_value" % ci] = last_value
    out["ncases"] = np.array(len(cases))
    out["gamma_tau"] = np.array([0.99, 0.95])
    np.savez_compressed(os.path.join(GOLDEN, "gae.npz"), **out)
    print("gae.npz:", len(out), "arrays")


def main():
    from oracle import reference_loader
    trl = reference_loader.load()
    os.makedirs(GOLDEN, exist_ok=True)
    only = sys.argv[1:]
    for name, fn in sorted(globals().items()):
        if name.startswith("golden_") and callable(fn):
            if only and name[len("golden_"):] not in only:
                continue
            fn(trl)


if __name__ == "__main__":
    main()
