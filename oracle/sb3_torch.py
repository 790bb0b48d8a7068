"""CPU ORACLE (test infrastructure, not product code): torch-eager restatement
of the stable-baselines3==1.7.0 arithmetic the reference drives.

SB3 is pinned by the reference (setup.py:17) but NOT vendored and not
installable here, so this file restates its published algorithm (SURVEY.md
Appendix A) with the same torch op sequence, cross-checked against the
reference's in-tree near copies:
  * ActorCriticPolicy construction / forward / evaluate_actions:
      pantheonrl/algos/modular/policies.py:84-118, 214-241, 273-290, 364-383
  * PPO.train:  pantheonrl/algos/adap/adap_learn.py:229-347 (minus context loss)
  * collect_rollouts: pantheonrl/algos/adap/adap_learn.py:415-471
PARITY: ppo_train reproduces the parameters and logged scalars of the reference's own
ADAP.train (adap_learn.py:229-347, executed verbatim with the context loss off) to 1e-7
(tests/golden/make_golden_sb3_intree.py, tests/test_oracle_sb3_intree.py).  UNPINNED: policy
construction / sampling against a real SB3 + torch 1.13.1.  This module is (a) the independent tolerance check of pth_oracle.c's hand-written
backward pass / Adam and of the CUDA update kernel, and (b) the CPU arm that
bench.py --impl reference times (torch CPU eager is what SB3 executes).
"""
import math

import numpy as np
import torch as th
from torch import nn


class MlpPolicy(nn.Module):
    """SB3 ActorCriticPolicy with MlpExtractor [dict(pi=[64,64], vf=[64,64])], Tanh."""

    def __init__(self, nvec=None, heads=(3,), box_dim=None, seed=None, lr=3e-4, adam_eps=1e-5, extra_inputs=0):
        super().__init__()
        if seed is not None:
            th.manual_seed(seed)  # SB3 set_random_seed (Appendix A1)
        self.nvec = None if nvec is None else [int(v) for v in nvec]
        self.heads = [int(h) for h in heads]
        # extra_inputs: AdapPolicy's context columns behind the features (adap/policies.py:71-84)
        F = (int(box_dim) if box_dim is not None else sum(self.nvec)) + int(extra_inputs)
        self.F, self.L = F, sum(self.heads)
        # creation order pi0, vf0, pi1, vf1 (MlpExtractor interleaves), then heads
        pi0, vf0 = nn.Linear(F, 64), nn.Linear(F, 64)
        pi1, vf1 = nn.Linear(64, 64), nn.Linear(64, 64)
        self.policy_net = nn.Sequential(pi0, nn.Tanh(), pi1, nn.Tanh())
        self.value_net_body = nn.Sequential(vf0, nn.Tanh(), vf1, nn.Tanh())
        self.action_net = nn.Linear(64, self.L)
        self.value_net = nn.Linear(64, 1)
        # orthogonal init, module by module (modular/policies.py:229-241)
        for mod, gain in ((self.policy_net, math.sqrt(2)), (self.value_net_body, math.sqrt(2)),
                          (self.action_net, 0.01), (self.value_net, 1.0)):
            for m in mod.modules():
                if isinstance(m, nn.Linear):
                    nn.init.orthogonal_(m.weight, gain=gain)
                    m.bias.data.fill_(0.0)
        # registration order of SB3: mlp_extractor.policy_net, .value_net, action_net, value_net
        self.optimizer = th.optim.Adam(self.ordered_parameters(), lr=lr, eps=adam_eps)

    def ordered_parameters(self):
        pi0, pi1 = self.policy_net[0], self.policy_net[2]
        vf0, vf1 = self.value_net_body[0], self.value_net_body[2]
        return [pi0.weight, pi0.bias, pi1.weight, pi1.bias, vf0.weight, vf0.bias, vf1.weight,
                vf1.bias, self.action_net.weight, self.action_net.bias, self.value_net.weight,
                self.value_net.bias]

    # ---- flat parameter vector in the engine's layout (first layers input-major)
    def to_flat(self):
        out = []
        for i, p in enumerate(self.ordered_parameters()):
            t = p.detach()
            if i in (0, 4):
                t = t.t()
            out.append(t.contiguous().reshape(-1))
        return th.cat(out).numpy().astype(np.float32)

    def from_flat(self, flat):
        flat = th.as_tensor(np.asarray(flat, np.float32))
        o = 0
        with th.no_grad():
            for i, p in enumerate(self.ordered_parameters()):
                n = p.numel()
                chunk = flat[o:o + n]
                if i in (0, 4):
                    p.copy_(chunk.reshape(p.shape[1], p.shape[0]).t())
                else:
                    p.copy_(chunk.reshape(p.shape))
                o += n
        return self

    def flat_grad(self):
        out = []
        for i, p in enumerate(self.ordered_parameters()):
            t = p.grad.detach()
            if i in (0, 4):
                t = t.t()
            out.append(t.contiguous().reshape(-1))
        return th.cat(out).numpy().astype(np.float32)

    # ---- SB3 preprocess_obs
    def features(self, obs):
        if self.nvec is None:
            return th.as_tensor(obs).float()
        obs = th.as_tensor(np.asarray(obs)).long()
        return th.cat([nn.functional.one_hot(obs[:, s], n).float() for s, n in enumerate(self.nvec)], dim=1)

    def _dists(self, latent_pi):
        logits = self.action_net(latent_pi)
        return logits, [th.distributions.Categorical(logits=lg) for lg in th.split(logits, self.heads, dim=1)]

    def evaluate_actions(self, obs, actions):
        x = self.features(obs)
        latent_pi, latent_vf = self.policy_net(x), self.value_net_body(x)
        _, dists = self._dists(latent_pi)
        actions = th.as_tensor(np.asarray(actions)).long()
        actions = actions.reshape(actions.shape[0], -1)  # Discrete: SB3 passes a flat vector
        log_prob = th.stack([d.log_prob(actions[:, h]) for h, d in enumerate(dists)], dim=1).sum(dim=1)
        entropy = th.stack([d.entropy() for d in dists], dim=1).sum(dim=1)
        return self.value_net(latent_vf), log_prob, entropy

    def forward_logits(self, obs):
        with th.no_grad():
            x = self.features(obs)
            logits = self.action_net(self.policy_net(x))
            values = self.value_net(self.value_net_body(x))
        return logits, values[:, 0]


def ppo_minibatch_step(policy, obs, actions, old_log_prob, advantages, returns, clip_range=0.2,
                       ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5, normalize_advantage=True):
    """One iteration of the inner loop of SB3 PPO.train (adap_learn.py:252-346)."""
    advantages = th.as_tensor(advantages).float()
    returns = th.as_tensor(returns).float()
    old_log_prob = th.as_tensor(old_log_prob).float()
    values, log_prob, entropy = policy.evaluate_actions(obs, actions)
    values = values.flatten()
    if normalize_advantage and len(advantages) > 1:
        advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    ratio = th.exp(log_prob - old_log_prob)
    policy_loss_1 = advantages * ratio
    policy_loss_2 = advantages * th.clamp(ratio, 1 - clip_range, 1 + clip_range)
    policy_loss = -th.min(policy_loss_1, policy_loss_2).mean()
    clip_fraction = th.mean((th.abs(ratio - 1) > clip_range).float()).item()
    value_loss = nn.functional.mse_loss(returns, values)
    entropy_loss = -th.mean(entropy)
    loss = policy_loss + ent_coef * entropy_loss + vf_coef * value_loss
    with th.no_grad():
        log_ratio = log_prob - old_log_prob
        approx_kl = th.mean((th.exp(log_ratio) - 1) - log_ratio).item()
    policy.optimizer.zero_grad()
    loss.backward()
    grad = policy.flat_grad()
    total_norm = th.nn.utils.clip_grad_norm_(policy.ordered_parameters(), max_grad_norm)
    policy.optimizer.step()
    return dict(pg_loss=policy_loss.item(), value_loss=value_loss.item(),
                entropy_loss=entropy_loss.item(), approx_kl=approx_kl, clip_fraction=clip_fraction,
                loss=loss.item(), grad_norm=float(total_norm), grad=grad)


def ppo_train(policy, obs, actions, old_log_prob, advantages, returns, perms, batch_size,
              **kw):
    """SB3 PPO.train over a flattened buffer: for each epoch's permutation
    (RolloutBuffer.get: indices = permutation(M); consecutive slices of
    batch_size; the last one may be short) run ppo_minibatch_step."""
    stats = []
    M = len(advantages)
    for perm in perms:
        perm = np.asarray(perm)
        for s in range(0, M, batch_size):
            idx = perm[s:s + batch_size]
            stats.append(ppo_minibatch_step(policy, obs[idx], actions[idx], old_log_prob[idx],
                                            advantages[idx], returns[idx], **kw))
    return stats


def bc_minibatch_step(policy, optimizer, obs, actions, ent_weight=1e-3, l2_weight=0.0):
    """One batch of behaviour cloning: BC._calculate_loss + the optimiser step of BC.train
    (pantheonrl/algos/bc.py:270-315, 345-352)."""
    _, log_prob, entropy = policy.evaluate_actions(obs, actions)
    prob_true_act = th.exp(log_prob).mean()
    log_prob, entropy = log_prob.mean(), entropy.mean()
    l2_norm = sum(th.sum(th.square(w)) for w in policy.ordered_parameters()) / 2
    ent_loss = -ent_weight * entropy
    neglogp = -log_prob
    loss = neglogp + ent_loss + l2_weight * l2_norm
    optimizer.zero_grad()
    loss.backward()
    grad = policy.flat_grad()
    optimizer.step()
    return dict(neglogp=neglogp.item(), loss=loss.item(), entropy=entropy.item(), ent_loss=ent_loss.item(),
                prob_true_act=prob_true_act.item(), l2_norm=l2_norm.item(), grad=grad)


def bc_train(policy, obs, actions, perms, batch_size=32, ent_weight=1e-3, l2_weight=0.0, lr=1e-3, eps=1e-8):
    """BC.train over epochs of shuffled batches with torch.optim.Adam at its defaults
    (bc.py:196: optimizer_cls=Adam, no optimizer_kwargs)."""
    opt = th.optim.Adam(policy.ordered_parameters(), lr=lr, eps=eps)
    stats, M = [], len(actions)
    for perm in perms:
        perm = np.asarray(perm)
        for s in range(0, M, batch_size):
            idx = perm[s:s + batch_size]
            stats.append(bc_minibatch_step(policy, opt, obs[idx], actions[idx], ent_weight, l2_weight))
    return stats


class AdapMlpPolicy(MlpPolicy):
    """AdapPolicy (pantheonrl/algos/adap/policies.py:21-131): the MlpPolicy whose two towers read
    cat(features, context); `context` is a [1, C] tensor the policy carries (set_context / get_context),
    and evaluate_actions takes rows of raw observation ++ the context stored with the sample."""

    def __init__(self, nvec=None, heads=(3,), box_dim=None, context_size=3, seed=None, **kw):
        super().__init__(nvec=nvec, heads=heads, box_dim=box_dim, seed=seed, extra_inputs=context_size, **kw)
        self.context_size = int(context_size)
        self.context = th.zeros(1, self.context_size)

    def set_context(self, ctxt):
        self.context = ctxt

    def get_context(self):
        return self.context

    def latent_pi(self, obs, context):
        """policy tower on `obs` rows with one context [1, C] for all of them (adap/policies.py:86-106)."""
        x = self.features(obs)
        x = th.cat((x, th.as_tensor(context).float().reshape(1, -1).repeat(x.shape[0], 1)), dim=1)
        return self.policy_net(x)

    def evaluate_actions(self, obs, actions):
        obs = th.as_tensor(np.asarray(obs)).float()
        x = th.cat((self.features(obs[:, :-self.context_size]), obs[:, -self.context_size:]), dim=1)
        latent_pi, latent_vf = self.policy_net(x), self.value_net_body(x)
        _, dists = self._dists(latent_pi)
        actions = th.as_tensor(np.asarray(actions)).long()
        actions = actions.reshape(actions.shape[0], -1)
        log_prob = th.stack([d.log_prob(actions[:, h]) for h, d in enumerate(dists)], dim=1).sum(dim=1)
        entropy = th.stack([d.entropy() for d in dists], dim=1).sum(dim=1)
        return self.value_net(latent_vf), log_prob, entropy


def adap_context_loss(policy, obs_rows, sidx, contexts):
    """get_context_kl_loss (pantheonrl/algos/adap/util.py:97-131) with the random draws handed in:
    `sidx` = th.randperm(B)[:num_state_samples], `contexts` [K, C] = the K sampled contexts.
    mean over the K (K - 1) / 2 context pairs of mean_states exp(-KL(dist_a || dist_b))."""
    from itertools import combinations
    states = th.as_tensor(np.asarray(obs_rows)).float()[:, :-policy.context_size][th.as_tensor(np.asarray(sidx)).long()]
    dists = [policy._dists(policy.latent_pi(states, c))[1] for c in th.as_tensor(np.asarray(contexts)).float()]
    kl = lambda a, b: sum(th.distributions.kl.kl_divergence(p, q) for p, q in zip(a, b))  # noqa: E731
    cls = [th.mean(th.exp(-kl(a, b))) for a, b in combinations(dists, 2)]
    return sum(cls) / len(cls)


def adap_minibatch_step(policy, obs, actions, old_log_prob, advantages, returns, sidx, contexts, context_loss_coeff=0.1,
                        clip_range=0.2, ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5):
    """One iteration of the inner loop of ADAP.train (adap_learn.py:252-346): PPO's losses + the context loss."""
    advantages = th.as_tensor(advantages).float()
    returns = th.as_tensor(returns).float()
    old_log_prob = th.as_tensor(old_log_prob).float()
    values, log_prob, entropy = policy.evaluate_actions(obs, actions)
    values = values.flatten()
    advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    ratio = th.exp(log_prob - old_log_prob)
    policy_loss = -th.min(advantages * ratio, advantages * th.clamp(ratio, 1 - clip_range, 1 + clip_range)).mean()
    value_loss = nn.functional.mse_loss(returns, values)
    entropy_loss = -th.mean(entropy)
    context_loss = adap_context_loss(policy, obs, sidx, contexts)
    loss = policy_loss + ent_coef * entropy_loss + vf_coef * value_loss + context_loss_coeff * context_loss
    policy.optimizer.zero_grad()
    loss.backward()
    grad = policy.flat_grad()
    total_norm = th.nn.utils.clip_grad_norm_(policy.ordered_parameters(), max_grad_norm)
    policy.optimizer.step()
    return dict(pg_loss=policy_loss.item(), value_loss=value_loss.item(), entropy_loss=entropy_loss.item(),
                context_loss=context_loss.item(), loss=loss.item(), grad_norm=float(total_norm), grad=grad)


def adap_train(policy, obs, actions, old_log_prob, advantages, returns, perms, batch_size, sidx, contexts, **kw):
    """ADAP.train over a flattened buffer; sidx [n_mb_total, S] (rows may be short: -1 padded), contexts
    [n_mb_total, K, C]."""
    stats, M, i = [], len(advantages), 0
    for perm in perms:
        perm = np.asarray(perm)
        for s0 in range(0, M, batch_size):
            idx = perm[s0:s0 + batch_size]
            si = np.asarray(sidx[i])
            stats.append(adap_minibatch_step(policy, obs[idx], actions[idx], old_log_prob[idx], advantages[idx],
                                             returns[idx], si[si >= 0][:len(idx)], contexts[i], **kw))
            i += 1
    return stats


class ModularMlpPolicy(MlpPolicy):
    """ModularPolicy (pantheonrl/algos/modular/policies.py:23-396) with its defaults: the main network is the
    MlpPolicy (two 64-64 tanh towers, action and value heads); every partner p owns a module that reads the
    main policy tower's latent: two more 64-64 tanh towers (`partner_mlp_extractor[p]`, BOTH fed with
    latent_pi), an action head and a value head.  logits = main + partner[p], value = main + partner[p]
    (:273-290, :307-326).  Construction / init order follows `_build` + `do_init_weights` (:229-270)."""

    def __init__(self, nvec=None, heads=(3,), box_dim=None, num_partners=1, seed=None, lr=3e-4, adam_eps=1e-5):
        super().__init__(nvec=nvec, heads=heads, box_dim=box_dim, seed=seed, lr=lr, adam_eps=adam_eps)
        self.num_partners = int(num_partners)
        mk = lambda: nn.Sequential(nn.Linear(64, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh())  # noqa: E731
        pp, pv, pa, pw = [], [], [], []
        for _ in range(self.num_partners):  # MlpExtractor interleaves pi / vf layer creation, then the two heads
            p0, v0, p1, v1 = nn.Linear(64, 64), nn.Linear(64, 64), nn.Linear(64, 64), nn.Linear(64, 64)
            pp.append(nn.Sequential(p0, nn.Tanh(), p1, nn.Tanh()))
            pv.append(nn.Sequential(v0, nn.Tanh(), v1, nn.Tanh()))
            pa.append(nn.Linear(64, self.L))
            pw.append(nn.Linear(64, 1))
        self.partner_policy_net, self.partner_value_body = nn.ModuleList(pp), nn.ModuleList(pv)
        self.partner_action_net, self.partner_value_net = nn.ModuleList(pa), nn.ModuleList(pw)
        # do_init_weights(init_main=True, init_partner=True): the main modules again, then per partner
        mods = [(self.policy_net, math.sqrt(2)), (self.value_net_body, math.sqrt(2)), (self.action_net, 0.01),
                (self.value_net, 1.0)]
        for p in range(self.num_partners):
            mods += [(self.partner_policy_net[p], math.sqrt(2)), (self.partner_value_body[p], math.sqrt(2)),
                     (self.partner_action_net[p], 0.01), (self.partner_value_net[p], 1.0)]
        for mod, gain in mods:
            for m in mod.modules():
                if isinstance(m, nn.Linear):
                    nn.init.orthogonal_(m.weight, gain=gain)
                    m.bias.data.fill_(0.0)
        self.optimizer = th.optim.Adam(self.ordered_parameters(), lr=lr, eps=adam_eps)

    def ordered_parameters(self):
        out = MlpPolicy.ordered_parameters(self)
        for p in range(getattr(self, "num_partners", 0)):
            pn, vb = self.partner_policy_net[p], self.partner_value_body[p]
            out += [pn[0].weight, pn[0].bias, pn[2].weight, pn[2].bias, vb[0].weight, vb[0].bias, vb[2].weight,
                    vb[2].bias, self.partner_action_net[p].weight, self.partner_action_net[p].bias,
                    self.partner_value_net[p].weight, self.partner_value_net[p].bias]
        return out

    def logits(self, obs, partner_idx):
        """(main_logits, partner_logits) — get_action_logits_from_obs (:385-396)."""
        latent_pi = self.policy_net(self.features(obs))
        return self.action_net(latent_pi), self.partner_action_net[partner_idx](self.partner_policy_net[partner_idx](latent_pi))

    def evaluate_actions(self, obs, actions, partner_idx=0):
        x = self.features(obs)
        latent_pi, latent_vf = self.policy_net(x), self.value_net_body(x)
        p_pi, p_vf = self.partner_policy_net[partner_idx](latent_pi), self.partner_value_body[partner_idx](latent_pi)
        logits = self.action_net(latent_pi) + self.partner_action_net[partner_idx](p_pi)
        dists = [th.distributions.Categorical(logits=lg) for lg in th.split(logits, self.heads, dim=1)]
        actions = th.as_tensor(np.asarray(actions)).long()
        actions = actions.reshape(actions.shape[0], -1)
        log_prob = th.stack([d.log_prob(actions[:, h]) for h, d in enumerate(dists)], dim=1).sum(dim=1)
        entropy = th.stack([d.entropy() for d in dists], dim=1).sum(dim=1)
        values = self.value_net(latent_vf) + self.partner_value_net[partner_idx](p_vf)
        return values, log_prob, entropy


def modular_minibatch_step(policy, partner_idx, obs, actions, old_log_prob, advantages, returns, marginal_reg_coef=0.0,
                           clip_range=0.2, ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5):
    """One iteration of the inner loop of ModularAlgorithm.train (modular/learn.py:246-326)."""
    advantages = th.as_tensor(advantages).float()
    returns = th.as_tensor(returns).float()
    old_log_prob = th.as_tensor(old_log_prob).float()
    values, log_prob, entropy = policy.evaluate_actions(obs, actions, partner_idx)
    values = values.flatten()
    advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    ratio = th.exp(log_prob - old_log_prob)
    policy_loss = -th.min(advantages * ratio, advantages * th.clamp(ratio, 1 - clip_range, 1 + clip_range)).mean()
    value_loss = nn.functional.mse_loss(returns, values)
    entropy_loss = -th.mean(entropy)
    # marginal regularisation (:298-318): softmax over ALL logits jointly, Wasserstein-1 with unit distances
    lg = [policy.logits(obs, idx) for idx in range(policy.num_partners)]
    main_logits = th.stack([m for m, _ in lg])
    composed_logits = main_logits + th.stack([p for _, p in lg])
    main_probs = th.mean(th.exp(main_logits - main_logits.logsumexp(dim=-1, keepdim=True)), dim=0)
    composed_probs = th.mean(th.exp(composed_logits - composed_logits.logsumexp(dim=-1, keepdim=True)), dim=0)
    marginal = th.mean(th.sum(th.abs(main_probs - composed_probs), dim=1))
    loss = policy_loss + ent_coef * entropy_loss + vf_coef * value_loss + marginal_reg_coef * marginal
    policy.optimizer.zero_grad()
    loss.backward()
    th.nn.utils.clip_grad_norm_(policy.ordered_parameters(), max_grad_norm)
    policy.optimizer.step()
    return dict(pg_loss=policy_loss.item(), value_loss=value_loss.item(), entropy_loss=entropy_loss.item(),
                marginal=marginal.item(), loss=loss.item())


def modular_train(policy, buffers, perms, batch_size, **kw):
    """ModularAlgorithm.train: for every partner, n_epochs passes over THAT partner's buffer
    (buffers[p] = (obs, actions, old_log_prob, advantages, returns), perms[p] = [n_epochs, M])."""
    stats = []
    for p, (obs, actions, old_log_prob, advantages, returns) in enumerate(buffers):
        M = len(advantages)
        for perm in perms[p]:
            perm = np.asarray(perm)
            for s0 in range(0, M, batch_size):
                idx = perm[s0:s0 + batch_size]
                stats.append(modular_minibatch_step(policy, p, obs[idx], actions[idx], old_log_prob[idx],
                                                    advantages[idx], returns[idx], **kw))
    return stats


class AdapMultPolicy(MlpPolicy):
    """AdapPolicyMult (pantheonrl/algos/adap/policies.py:134-283): each tower is
    x = tanh(W0 f + b0); s = tanh(Ws x + bs) (64 -> 64 C); y_j = x_j + sum_c s[j C + c] ctx_c; out = tanh(W1 y + b1)
    on the features WITHOUT the context (MultModel.forward :263-267).  Module creation order follows MultModel.__init__:
    pi0, vf0, pi1, vf1 (zip_longest), agent_scaling, value_scaling, then the heads."""

    def __init__(self, nvec=None, heads=(3,), box_dim=None, context_size=3, seed=None, lr=3e-4, adam_eps=1e-5):
        nn.Module.__init__(self)
        if seed is not None:
            th.manual_seed(seed)
        self.nvec = None if nvec is None else [int(v) for v in nvec]
        self.heads = [int(h) for h in heads]
        F = int(box_dim) if box_dim is not None else sum(self.nvec)
        self.F, self.L, self.context_size = F, sum(self.heads), int(context_size)
        C = self.context_size
        pi0, vf0 = nn.Linear(F, 64), nn.Linear(F, 64)
        pi1, vf1 = nn.Linear(64, 64), nn.Linear(64, 64)
        self.agent_branch_1 = nn.Sequential(pi0, nn.Tanh())
        self.agent_scaling = nn.Sequential(nn.Linear(64, 64 * C), nn.Tanh())
        self.agent_branch_2 = nn.Sequential(pi1, nn.Tanh())
        self.value_branch_1 = nn.Sequential(vf0, nn.Tanh())
        self.value_scaling = nn.Sequential(nn.Linear(64, 64 * C), nn.Tanh())
        self.value_branch_2 = nn.Sequential(vf1, nn.Tanh())
        self.action_net = nn.Linear(64, self.L)
        self.value_net = nn.Linear(64, 1)
        self.context = th.zeros(1, C)
        # ActorCriticPolicy._build: orthogonal init module by module (mlp_extractor = the MultModel: gain sqrt 2)
        for mod, gain in ((self.agent_branch_1, math.sqrt(2)), (self.agent_scaling, math.sqrt(2)),
                          (self.agent_branch_2, math.sqrt(2)), (self.value_branch_1, math.sqrt(2)),
                          (self.value_scaling, math.sqrt(2)), (self.value_branch_2, math.sqrt(2)),
                          (self.action_net, 0.01), (self.value_net, 1.0)):
            for m in mod.modules():
                if isinstance(m, nn.Linear):
                    nn.init.orthogonal_(m.weight, gain=gain)
                    m.bias.data.fill_(0.0)
        self.optimizer = th.optim.Adam(self.ordered_parameters(), lr=lr, eps=adam_eps)

    def ordered_parameters(self):
        a1, as_, a2 = self.agent_branch_1[0], self.agent_scaling[0], self.agent_branch_2[0]
        v1, vs, v2 = self.value_branch_1[0], self.value_scaling[0], self.value_branch_2[0]
        return [a1.weight, a1.bias, as_.weight, as_.bias, a2.weight, a2.bias, v1.weight, v1.bias, vs.weight, vs.bias,
                v2.weight, v2.bias, self.action_net.weight, self.action_net.bias, self.value_net.weight,
                self.value_net.bias]

    _TRANSPOSED = (0, 6)  # the two first-layer matrices are stored input-major in the flat vector

    def to_flat(self):
        out = []
        for i, p in enumerate(self.ordered_parameters()):
            t = p.detach()
            out.append((t.t() if i in self._TRANSPOSED else t).contiguous().reshape(-1))
        return th.cat(out).numpy().astype(np.float32)

    def from_flat(self, flat):
        flat = th.as_tensor(np.asarray(flat, np.float32))
        o = 0
        with th.no_grad():
            for i, p in enumerate(self.ordered_parameters()):
                n = p.numel()
                chunk = flat[o:o + n]
                p.copy_(chunk.reshape(p.shape[1], p.shape[0]).t() if i in self._TRANSPOSED else chunk.reshape(p.shape))
                o += n
        return self

    def set_context(self, ctxt):
        self.context = ctxt

    def get_context(self):
        return self.context

    def _tower(self, b1, sc, b2, feats, ctx):
        x = b1(feats)
        xa = sc(x).view(feats.shape[0], 64, self.context_size)
        return b2(x + th.matmul(xa, ctx.unsqueeze(-1)).squeeze(-1))

    def latent_pi(self, obs, context):
        f = self.features(obs)
        ctx = th.as_tensor(context).float().reshape(1, -1).repeat(f.shape[0], 1)
        return self._tower(self.agent_branch_1, self.agent_scaling, self.agent_branch_2, f, ctx)

    def evaluate_actions(self, obs, actions):
        obs = th.as_tensor(np.asarray(obs)).float()
        f, ctx = self.features(obs[:, :-self.context_size]), obs[:, -self.context_size:]
        latent_pi = self._tower(self.agent_branch_1, self.agent_scaling, self.agent_branch_2, f, ctx)
        latent_vf = self._tower(self.value_branch_1, self.value_scaling, self.value_branch_2, f, ctx)
        _, dists = self._dists(latent_pi)
        actions = th.as_tensor(np.asarray(actions)).long()
        actions = actions.reshape(actions.shape[0], -1)
        log_prob = th.stack([d.log_prob(actions[:, h]) for h, d in enumerate(dists)], dim=1).sum(dim=1)
        entropy = th.stack([d.entropy() for d in dists], dim=1).sum(dim=1)
        return self.value_net(latent_vf), log_prob, entropy
