"""RL stack: environment physics/observations, PPO problem API, consensus PPO trainers."""
import numpy as np
import networkx as nx
import pytest
import torch

from nn_distributed_training_b200.rl import (DSGDPPO, DSGTPPO, DiNNOPPO, DistPPOProblem, FFReLUNet, PPO, SimpleTagEnv,
                                              heuristic_prey_action)


def _env(E=4, steps=12, seed=0):
    return SimpleTagEnv(num_envs=E, num_good=1, num_adversaries=3, num_obstacles=8, max_cycles=steps, seed=seed)


def test_env_shapes_and_spaces():
    env = _env()
    assert env.observation_spaces["adversary_0"].shape == (12,)      # reference: obs dim 12, act dim 5
    assert env.observation_spaces["agent_0"].shape == (10,)
    assert env.action_spaces["adversary_0"].shape == (5,)
    oa, og = env.reset()
    assert oa.shape == (4, 3, 12) and og.shape == (4, 1, 10)
    # own position is obs[2:4]; relative position of another agent is pos_other - pos_self
    torch.testing.assert_close(oa[:, 0, 2:4], env.pos[:, 0])
    torch.testing.assert_close(oa[:, 0, 4:6], env.pos[:, 1] - env.pos[:, 0])
    torch.testing.assert_close(oa[:, 0, 10:12], env.vel[:, 3])        # prey velocity


def test_env_physics_matches_scalar_reference_step():
    """One world stepped by an independent scalar implementation of the MPE integrator."""
    env = _env(E=1, steps=5, seed=3)
    env.reset()
    pos0, vel0 = env.pos[0].double().numpy().copy(), env.vel[0].double().numpy().copy()
    act = torch.rand(1, 4, 5)
    env.step(act)
    a = act[0].double().numpy()
    size = np.array([0.075] * 3 + [0.05]); accel = np.array([3.0] * 3 + [4.0]); vmax = np.array([1.0] * 3 + [1.3])
    obst = np.array([(-1.2, -0.6), (0.1, -1.1), (-0.3, 0.4), (0.9, 0.75), (-0.9, 1.2), (-0.1, 1.3), (-1.2, 0.0), (1.3, 0.0)])
    force = np.stack([a[:, 1] - a[:, 2], a[:, 3] - a[:, 4]], 1) * accel[:, None]
    def coll(pa, pb, dmin):
        d = pa - pb; dist = np.sqrt((d ** 2).sum()); k = 1e-3
        return 1e2 * d / dist * np.logaddexp(0, -(dist - dmin) / k) * k
    for i in range(4):
        for j in range(4):
            if i != j:
                force[i] += coll(pos0[i], pos0[j], size[i] + size[j])
        for o in obst:
            force[i] += coll(pos0[i], o, size[i] + 0.2)
    vel = vel0 * 0.75 + force * 0.1
    for i in range(4):
        s = np.linalg.norm(vel[i])
        if s > vmax[i]:
            vel[i] = vel[i] / s * vmax[i]
    np.testing.assert_allclose(env.vel[0].numpy(), vel, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(env.pos[0].numpy(), pos0 + vel * 0.1, rtol=1e-4, atol=1e-5)


def test_rewards_and_heuristic():
    env = _env(E=2, steps=5, seed=1)
    env.reset()
    env.pos[0, :, :] = torch.tensor([[0.0, 0.0], [0.5, 0.5], [-0.5, 0.5], [0.05, 0.0]])   # predator 0 touches the prey
    env.pos[1, :, :] = torch.tensor([[0.0, 0.0], [0.5, 0.5], [-0.5, 0.5], [0.95, 0.0]])
    r_adv, r_good = env._rewards()
    d0 = torch.tensor([0.05, (0.45 ** 2 + 0.5 ** 2) ** 0.5, (0.55 ** 2 + 0.5 ** 2) ** 0.5])
    assert r_adv[0, 0].item() == pytest.approx(10.0 - 0.1 * d0.sum().item(), abs=1e-4)
    assert (r_adv[0] == r_adv[0, 0]).all()                      # shared among predators
    assert r_good[0, 0].item() == pytest.approx(-10.0, abs=1e-5)
    assert r_good[1, 0].item() == pytest.approx(-0.5, abs=1e-4)   # boundary penalty (0.95-0.9)*10
    _, og = env.observe()
    act = heuristic_prey_action(og[:, 0], 3)
    assert act.shape == (2, 5) and (act >= 0).all() and act[0, 1] > 0    # flee +x from predator at -x


def _problem(E=4, steps=12, tpb=60):
    env = _env(E=E, steps=steps)
    obs_dim, act_dim = 12, 5
    torch.manual_seed(0)
    return DistPPOProblem(FFReLUNet([obs_dim, 16, 16, act_dim]), FFReLUNet([obs_dim, 16, 16, 1]), nx.wheel_graph(3), env,
                          timesteps_per_batch=tpb, max_timesteps_per_episode=steps * 4, gamma=0.99,
                          n_updates_per_iteration=2, lr=3e-4, clip=0.2, save_freq=1)


def test_rollout_rtgs_and_losses():
    pr = _problem()
    pr.split_rollout_marl()
    n = pr.curr_obs[0].shape[0]
    assert pr.curr_acts[0].shape == (n, 5) and pr.curr_log_probs[0].shape == (n,) and pr.curr_rtgs[0].shape == (n,)
    assert n * 3 >= 60 and pr.logger["t_so_far"] > 0
    # rewards-to-go agree with the reference's scalar recursion
    rt = pr.compute_rtgs([[1.0, 2.0, 3.0]])
    assert rt.tolist() == pytest.approx([1 + 0.99 * (2 + 0.99 * 3), 2 + 0.99 * 3, 3.0])
    pr.update_advantage()
    assert abs(pr.A_k[0].mean().item()) < 1e-5
    a, c = pr.ev_ppo_loss(1)
    assert torch.isfinite(a) and torch.isfinite(c) and abs(a.item()) < 1e-4   # ratio = 1 at the sampling policy
    assert pr.local_batch_loss(1).requires_grad


@pytest.mark.parametrize("cls,conf", [
    (DiNNOPPO, {"rho_init": 1.0, "rho_scaling": 1.0, "primal_lr_start": 3e-4, "primal_lr_finish": 1e-3,
                "lr_decay_type": "constant", "persistant_primal_opt": False, "primal_iterations": 2,
                "max_rl_timesteps": 400, "outer_iterations": 10_000_000, "ID": 1}),
    (DSGDPPO, {"alpha0": 3e-4, "mu": 0.0, "max_rl_timesteps": 400, "ID": 1}),
    (DSGTPPO, {"alpha_actor": 3e-4, "alpha_critic": 1e-3, "max_rl_timesteps": 400, "ID": 1}),
])
def test_consensus_ppo_trainers(tmp_path, cls, conf):
    pr = _problem()
    before = torch.nn.utils.parameters_to_vector(pr.models[0].actor.parameters()).clone()
    tr = cls(pr, "cpu", dict(conf, out_dir=str(tmp_path)))
    tr.train()
    after = torch.nn.utils.parameters_to_vector(pr.models[0].actor.parameters())
    assert not torch.equal(before, after) and torch.isfinite(after).all()
    assert len(tr.avg_ep_rews) >= 2 and len(tr.agreements[0]) == 3
    alg = tr.alg
    files = {f.name for f in tmp_path.iterdir()}
    assert f"avg_ep_rews_{alg}_1.npy" in files and f"agreements_{alg}_1.npz" in files
    assert any(f.startswith(f"ppo_actors_tag_{alg}_1_") for f in files)
    sd = torch.load(tmp_path / f"ppo_actors_tag_{alg}_1_0.pth", weights_only=False)
    assert set(sd) == {"actor0", "actor1", "actor2"}


def test_dsgt_per_slot_alpha():
    pr = _problem()
    tr = DSGTPPO(pr, "cpu", {"alpha_actor": 1e-3, "alpha_critic": 5e-3, "max_rl_timesteps": 100, "ID": 2, "writeout": False})
    slots = {s.name: s for s in tr.inner.arena.layout.slots}
    a = tr.inner.alpha
    s_a = next(s for n, s in slots.items() if n.startswith("actor"))
    s_c = next(s for n, s in slots.items() if n.startswith("critic"))
    assert a[s_a.offset].item() == pytest.approx(1e-3) and a[s_c.offset].item() == pytest.approx(5e-3)


def test_centralized_ppo_and_eval(tmp_path):
    from nn_distributed_training_b200.rl import main as rl_main
    env = _env(E=4, steps=10)
    m = PPO(FFReLUNet, env, timesteps_per_batch=60, max_timesteps_per_episode=40, n_updates_per_iteration=2, lr=3e-4,
            save_freq=1, out_dir=str(tmp_path), ID=7)
    m.learn(total_timesteps=300)
    assert (tmp_path / "ppo_actor_tag_7.pth").exists() and len(m.avg_ep_rews) >= 1
    ret = rl_main.test(env, str(tmp_path / "ppo_actor_tag_7.pth"))
    assert np.isfinite(ret)


def test_env_render_and_rollout_gif(tmp_path):
    """Rendering (reference: pyglet window / vids/*.mp4): PIL frames of a world and a GIF of a recorded rollout;
    ``render=True`` in the trainers writes one every ``render_every_i`` iterations."""
    import os
    from nn_distributed_training_b200.rl.eval_policy import rollout, save_rollout_gif
    from nn_distributed_training_b200.rl.simple_tag import SimpleTagEnv
    from nn_distributed_training_b200.rl.model import FFReLUNet
    env = SimpleTagEnv(num_envs=2, num_obstacles=8, max_cycles=6, seed=0)
    im = env.render(size=120)
    assert im.size == (120, 120) and len(im.getcolors(maxcolors=1 << 16)) >= 4        # background, obstacles, predators, prey
    torch.manual_seed(0)
    actor = FFReLUNet([env.observation_spaces["adversary_0"].shape[0], 16, 5])
    ret, length, traj = rollout(actor, env, record=True)
    assert traj.shape == (6, 4, 2) and length == 6
    out = save_rollout_gif(env, traj, os.path.join(str(tmp_path), "ep.gif"), size=100)
    assert os.path.getsize(out) > 200
    from nn_distributed_training_b200.rl.ppo import PPO
    agent = PPO(FFReLUNet, SimpleTagEnv(num_envs=2, num_obstacles=2, max_cycles=5, seed=1), timesteps_per_batch=20,
                max_timesteps_per_episode=5, n_updates_per_iteration=1, render=True, render_every_i=1, save_freq=100,
                out_dir=str(tmp_path), seed=0)
    agent.learn(total_timesteps=20)
    assert any(f.startswith("render_0_") and f.endswith(".gif") for f in os.listdir(str(tmp_path)))


def test_reward_and_agreement_plots_read_the_reference_file_names(tmp_path):
    """RL/plot_reward.py / plot_agreements.py equivalents: centralized ``avg_ep_rews_<ID>.npy``, distributed
    ``avg_ep_rews_<alg>_<ID>.npy`` (``cadmm`` = DiNNO), mean + min/max band over runs, PNG without matplotlib."""
    import os
    from nn_distributed_training_b200.rl import plot_reward as pr
    d = str(tmp_path)
    t = np.arange(1, 21) * 100
    for ID in (1, 2, 3):
        np.save(os.path.join(d, f"timesteps_{ID}.npy"), t)
        np.save(os.path.join(d, f"avg_ep_rews_{ID}.npy"), np.linspace(-30, 10 + ID, 20))
    for alg, ID in (("cadmm", 0), ("dinno", 7), ("dsgt", 13), ("dsgd", 4)):
        np.save(os.path.join(d, f"timesteps_{alg}_{ID}.npy"), t[:15])
        np.save(os.path.join(d, f"avg_ep_rews_{alg}_{ID}.npy"), np.linspace(-40, 5 + ID, 15))
    np.savez(os.path.join(d, "agreements_dsgt_13"), agree_0=np.geomspace(1, 1e-3, 15), agree_1=np.geomspace(1, 2e-3, 15),
             agree_2=np.geomspace(1, 3e-3, 15))
    s = pr.summarize(d)
    assert s["centralized"]["runs"] == 3 and s["dinno"]["runs"] == 2 and s["dsgt"]["runs"] == 1 and s["dsgd"]["runs"] == 1
    assert abs(s["centralized"]["final_mean"] - 12.0) < 1e-9
    out = pr.plot(d, out=os.path.join(d, "RL_reward.png"))
    assert os.path.getsize(out) > 500
    out = pr.plot_agreements(os.path.join(d, "agreements_dsgt_13.npz"), out=os.path.join(d, "RL_agreement.png"))
    assert os.path.getsize(out) > 500


def test_train_script_cli_writes_the_artefacts(tmp_path):
    """`python -m nn_distributed_training_b200.rl.train_dsgd_multi ...` (reference: RL/dist_rl/train_dsgd_multi.py):
    one iteration, files named as the reference names them, episode GIF when --render."""
    import os
    from nn_distributed_training_b200.rl import plot_reward, train_dsgd_multi
    out = str(tmp_path / "trained")
    train_dsgd_multi.main(["--max_rl_timesteps", "1500", "--num_envs", "4", "--out_dir", out, "--save_freq", "1",
                           "--render", "--render_every_i", "1", "--seed", "0", "--ID", "3"])
    files = set(os.listdir(out))
    assert {"ppo_actors_tag_dsgd_3_0.pth", "ppo_critics_tag_dsgd_3_0.pth", "avg_ep_rews_dsgd_3.npy", "timesteps_dsgd_3.npy",
            "agreements_dsgd_3.npz", "dsgd_3_0.gif"} <= files
    s = plot_reward.summarize(out)
    assert s["dsgd"]["runs"] == 1


def test_shipped_dinno_ppo_policies_catch_the_prey():
    """R10: the trained artefacts shipped under rl/trained (reference file names) load through the evaluation tool and the
    DiNNO-PPO predators score far above the untrained plateau (-90) — see profiles/rl_rewards.md."""
    import os
    import numpy as np
    from nn_distributed_training_b200.rl.eval_policy import load_distributed_actors, rollout
    from nn_distributed_training_b200.rl.model import FFReLUNet
    from nn_distributed_training_b200.rl.simple_tag import SimpleTagEnv
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nn_distributed_training_b200", "rl", "trained")
    curve = np.load(os.path.join(d, "avg_ep_rews_dinno_0.npy"))
    assert curve[:10].mean() < -50 and curve[-50:].mean() > 380          # the reference's DiNNO-PPO ends at 380-498
    assert np.load(os.path.join(d, "avg_ep_rews_dsgd_0.npy"))[-50:].mean() < 0      # DSGD-PPO never learns (as in the reference)
    env = SimpleTagEnv(num_envs=16, num_good=1, num_adversaries=3, num_obstacles=8, max_cycles=50, device="cpu", seed=5)
    actors = load_distributed_actors(os.path.join(d, "ppo_actors_tag_dinno_0_3000.pth"), lambda: FFReLUNet([12, 64, 64, 64, 5]), 3)
    ret, length, _ = rollout(actors, env)
    assert length == 50 and ret.mean() > 300
