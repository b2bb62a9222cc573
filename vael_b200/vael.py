"""VAEL models: host-side mirror of models/vael.py with the logic path on CUDA kernels.

Same classes, constructor signatures, forward signatures and side-effect attributes as the
reference (`facts_probs`, `worlds_prob`, `query_prob`, `world`, `z_subsym`, `semiring`,
`model_dict`, `herbrand_base`, `w_q`, `W_*`), so the reference's train / metrics code drives
it unchanged.  What happens between the MLP and the decoder is different:

  reference (models/vael.py)                               here
  :159-177  softmax/eps/renorm, 8 ATen kernels             facts_from_logits      (1 kernel)
  :200-225  outer product + masked sum                     circuit_forward        (1 kernel, TMA-tiled)
  :98-135   python SDD walk, ~3k semiring ops              CompiledProgram.evaluate (same kernel)
  :61-65    log + gumbel_softmax + matmul, ~9 kernels      gumbel_world           (1 kernel)
  autograd through all of it                               one adjoint kernel per stage

The encoder / decoder stay PyTorch modules passed in by the caller.
"""
from __future__ import annotations

import torch
from torch import nn

from .graph_semiring import GraphSemiring
from .logic import Constant, Term
from .ops import (circuit_forward, circuit_queries_all, facts_from_logits, facts_from_logits_clipped, fused_logic_supported,
                  gumbel_world, logic_train)
from .program import CompiledProgramSet, EvaluationResult


class VAELModel(nn.Module):
    def __init__(self, encoder, decoder, mlp, latent_dims, model_dict, dropout=None, is_train=True, device="cpu"):
        super().__init__()
        self.encoder, self.mlp, self.decoder = encoder, mlp, decoder
        self.dropout = dropout
        self.latent_dim_sym, self.latent_dim_sub = latent_dims
        self.model_dict = model_dict  # {'query': {label: CompiledProgram}, 'evidence': {...}}
        self.is_train = is_train
        self.device = device
        self.herbrand_base = self.define_herbrand_base(self.mlp.n_facts).to(self.device)
        self.weights_dict = self.build_weights_dictionary(self.mlp.n_facts)
        self.gumbel_noise = None     # tests inject [B,W] Gumbel noise here (SURVEY.md hard part 7)
        self.reparam_noise = None    # and [B,L] normal noise here
        self._world_cache = None
        # True: Gumbel noise from torch's (CUDA-graph-safe) device generator instead of the kernel's own
        # Philox stream seeded from the host — needed when the step is captured in a CUDA graph
        self.graph_safe_rng = False

    # ---- sampled world: kept as (index, y_soft), materialised one-hot on demand -----------------
    @property
    def world(self):
        if isinstance(self._world_cache, tuple) and isinstance(self._world_cache[0], str):   # fused forward: index only
            _, idx, n = self._world_cache
            self._world_cache = torch.zeros(idx.shape[0], n * n, device=idx.device).scatter_(1, idx.long()[:, None], 1.0)
        elif isinstance(self._world_cache, tuple):
            idx, y_soft = self._world_cache
            hard = torch.zeros_like(y_soft).scatter_(1, idx.long()[:, None], 1.0)
            self._world_cache = hard - y_soft.detach() + y_soft
        return self._world_cache

    @world.setter
    def world(self, value):
        self._world_cache = value

    def decode(self, z, w):
        """z: (bs, latent_dim_sub), w: (bs, herbrand_base)"""
        return self.decoder(torch.cat([z, w], 1))

    def sample_world(self, scores, herbrand, is_log=False):
        """Gumbel-softmax(hard) world + its Herbrand vector in one kernel (models/vael.py:61-65)."""
        noise = self.gumbel_noise
        if noise is None and self.graph_safe_rng:
            noise = -torch.empty_like(scores).exponential_().log()
        world_h, idx, y_soft = gumbel_world(scores, herbrand, tau=1.0, is_log=is_log, noise=noise)
        self._world_cache = (idx, y_soft)
        self.world_idx = idx
        return world_h

    def forward(self, x, labels):
        mu, logvar = self.encoder(x)
        return self._forward_after_encoder(mu, logvar, labels)

    def _forward_after_encoder(self, mu, logvar, labels):
        z = self.reparametrize(mu, logvar)
        self.z_subsym = z[:, self.latent_dim_sym:]
        if self.dropout and self.training:
            self.z_subsym = nn.functional.dropout(self.z_subsym, p=self.dropout)
        self.facts_probs = self.compute_facts_probability(z[:, :self.latent_dim_sym])
        self.query_prob, self.worlds_prob = self.problog_inference(self.facts_probs, labels=labels)
        world_h = self.sample_world(self.worlds_prob, self.herbrand_base)
        image = self.decode(self.z_subsym, world_h)
        return image, mu, logvar, self.query_prob

    def reparametrize(self, mu, logvar):
        """z = mu + exp(logvar/2) * eps; gated on the constructor flag like the reference
        (models/vael.py:72-80), noise drawn on the device instead of on the host."""
        if not self.is_train:
            return mu
        eps = self.reparam_noise if self.reparam_noise is not None else torch.randn_like(mu)
        return mu + torch.exp(logvar / 2) * eps

    def herbrand(self, world):
        return torch.matmul(world, self.herbrand_base)

    def define_herbrand_base(self, n_facts):
        raise NotImplementedError

    def build_weights_dictionary(self, n_facts):
        raise NotImplementedError

    def compute_facts_probability(self, z):
        raise NotImplementedError

    def update_semiring_weights(self, facts_probs):
        raise NotImplementedError

    def extract_worlds_probability(self, res):
        raise NotImplementedError

    def extract_query_probability(self, res):
        raise NotImplementedError

    # ---- generic compiled-program path (models/vael.py:98-135) --------------------------------------
    def problog_inference(self, facts_probs, labels, query=None):
        """Query-mode evaluation of the compiled program selected by `query` (an int label)."""
        if query is None:
            query = int(labels[0][-2]) if torch.is_tensor(labels) and labels.dim() == 2 else int(labels)
        self.update_semiring_weights(facts_probs)
        res = self.model_dict["query"][query].evaluate(semiring=self.semiring)
        query_prob = self.extract_query_probability(res)
        self.worlds_prob = self.extract_worlds_probability(res)
        return query_prob, self.worlds_prob

    def problog_inference_with_evidence(self, facts_probs, evidence):
        """P(w | evidence) from the evidence-mode program (models/vael.py:119-135)."""
        self.update_semiring_weights(facts_probs)
        res = self.model_dict["evidence"][int(evidence)].evaluate(semiring=self.semiring)
        return self.extract_worlds_probability(res)


class MNISTPairsVAELModel(VAELModel):
    def __init__(self, encoder, decoder, mlp, latent_dims, model_dict, w_q, dropout=None, is_train=True, device=False,
                 query_per_row: bool = False):
        super().__init__(encoder=encoder, decoder=decoder, mlp=mlp, latent_dims=latent_dims, model_dict=model_dict,
                         dropout=dropout, is_train=is_train, device=device)
        self.w_q = w_q
        # False = the reference's behaviour: one query per batch taken from labels[0] (models/vael.py:212-213);
        # True = one query per row from labels[:, -2] (mixed-label synthetic batches, no host sync)
        self.query_per_row = query_per_row
        self._pset = None
        # forward() runs the logic path (facts head -> circuit -> Gumbel sampling -> Herbrand vector) as ONE kernel
        # per direction when the compiled program allows it; False = the stage-by-stage kernels
        self.fused_logic = True
        self._worlds_prob = None

    # worlds_prob: the reference sets it in problog_inference and nobody reads it during training; the fused
    # forward never materialises it, so it is computed from facts_probs on first access
    @property
    def worlds_prob(self):
        if self._worlds_prob is None and getattr(self, "facts_probs", None) is not None:
            dc = self.program_set().device_circuit(self.facts_probs.device)
            self._worlds_prob, _ = circuit_forward(dc, self.facts_probs, None)
        return self._worlds_prob

    @worlds_prob.setter
    def worlds_prob(self, value):
        self._worlds_prob = value

    def mlp_logits(self, z):
        return self.mlp.logits(z) if hasattr(self.mlp, "logits") else torch.cat(self.mlp(z), dim=1)

    def forward(self, x, labels):
        """models/vael.py:39-70.  With `fused_logic` the four logic stages are one kernel (ops.logic_train); the
        side-effect attributes keep their meaning: facts_probs, query_prob, world_idx are written, worlds_prob and
        world (the hard one-hot; its straight-through soft part lives inside the fused adjoint) on first access."""
        if not self.fused_logic:
            return super().forward(x, labels)
        mu, logvar = self.encoder(x)
        dc = self.program_set().device_circuit(mu.device)
        if not fused_logic_supported(dc):
            self.fused_logic = False
            return self._forward_after_encoder(mu, logvar, labels)
        z = self.reparametrize(mu, logvar)
        self.z_subsym = z[:, self.latent_dim_sym:]
        if self.dropout and self.training:
            self.z_subsym = nn.functional.dropout(self.z_subsym, p=self.dropout)
        logits = self.mlp_logits(z[:, :self.latent_dim_sym])
        query = labels[:, -2] if self.query_per_row else int(labels[0][-2])
        seed_dev = None
        if self.gumbel_noise is None and self.graph_safe_rng:
            seed_dev = torch.randint(0, 2 ** 62, (1,), device=logits.device, dtype=torch.int64)
        facts, self.query_prob, world_h, idx, _ = logic_train(dc, logits, query, eps=1e-5, tau=1.0,
                                                              noise=self.gumbel_noise, seed_dev=seed_dev)
        self.facts_probs = facts.reshape(-1, 2, self.mlp.n_facts // 2)
        self._worlds_prob = None
        self.world_idx = idx
        self._world_cache = ("hard", idx, self.mlp.n_facts // 2)
        image = self.decode(self.z_subsym, world_h)
        return image, mu, logvar, self.query_prob

    def program_set(self) -> CompiledProgramSet:
        """The compiled addition family behind `problog_inference`.  The reference computes the query from the
        caller's `w_q` matrix (models/vael.py:200-206); here it comes from the compiled circuit, so on first use
        `w_q` is checked against the circuit's own world -> query table: a custom or permuted matrix raises
        instead of being silently ignored.  Extra all-zero columns of `w_q` (the reference's column 19,
        mnist_task_VAEL.py:226-228) are remembered: such a query is 0 for every row."""
        if self._pset is None:
            any_view = next(iter(self.model_dict["query"].values())) if self.model_dict else None
            if any_view is not None and hasattr(any_view, "pset") and any_view.pset.circuit.n_queries > 1:
                pset = any_view.pset
            else:  # no model_dict (the reference's training loop never reads it for MNIST), or one compiled
                #    program per label (the reference's own build_model_dict through vael_b200.compat)
                from .mnist_task import compile_addition_family
                pset = CompiledProgramSet(compile_addition_family(2, self.mlp.n_facts // 2))
            if self.w_q is not None:
                mine = torch.from_numpy(pset.circuit.worlds_queries_matrix()).to(torch.float32)
                given = self.w_q.detach().to("cpu", torch.float32)
                nq = mine.shape[1]
                if given.shape[0] != mine.shape[0] or given.shape[1] < nq or not torch.equal(given[:, :nq], mine) \
                        or bool(given[:, nq:].any()):
                    raise RuntimeError("vael_b200: the worlds-queries matrix w_q passed to the model differs from the "
                                       "compiled program's (build it with build_worlds_queries_matrix)")
                self._n_wq_cols = given.shape[1]
            self._pset = pset
        return self._pset

    def compute_facts_probability(self, z):
        """[B, latent_dim_sym] -> [B, 2, n]: MLP, then softmax / +1e-5 / detached renorm per digit
        (models/vael.py:159-177) in one kernel."""
        n = self.mlp.n_facts // 2
        return facts_from_logits(self.mlp_logits(z), n_groups=2, eps=1e-5).reshape(-1, 2, n)

    def define_herbrand_base(self, n_facts):
        """Row i*n+k = [onehot(i), onehot(k)] (models/vael.py:179-187)."""
        n = n_facts // 2
        eye = torch.eye(n)
        return torch.cat([eye.repeat_interleave(n, dim=0), eye.repeat(n, 1)], dim=1)

    def build_weights_dictionary(self, n_facts):
        """Term('p_{pos}{digit}') keys in fact-column order (models/vael.py:189-198)."""
        return {Term(f"p_{pos}{d}"): "NOT DEFINED" for pos in (1, 2) for d in range(n_facts // 2)}

    def compute_query(self, query, worlds_prob):
        """sum_w w_q[w, query] * P(w)  (models/vael.py:200-206); kept for callers holding worlds."""
        return torch.sum(self.w_q[:, query] * worlds_prob, dim=1, keepdim=True)

    def problog_inference(self, facts_probs, labels=None, query=None):
        """Worlds and query probability from the compiled circuit — observably the dense override
        of the reference (models/vael.py:208-225): raw, un-normalised products."""
        if query is None:
            query = labels[:, -2] if self.query_per_row else int(labels[0][-2])
        dc = self.program_set().device_circuit(facts_probs.device)
        if not torch.is_tensor(query) and dc.n_queries <= int(query) < getattr(self, "_n_wq_cols", 0):
            # an all-zero column of w_q (sum 19 of two digits): no world satisfies it
            self.worlds_prob, _ = circuit_forward(dc, facts_probs, None)
            return torch.zeros(facts_probs.shape[0], 1, device=facts_probs.device), self.worlds_prob
        self.worlds_prob, query_prob = circuit_forward(dc, facts_probs, query)
        return query_prob, self.worlds_prob

    def all_queries(self, facts_probs):
        """[B, 19] probabilities of every sum at once (replaces the 19-call loop of
        utils/mnist_utils/metrics.py:71-75)."""
        return circuit_queries_all(self.program_set().device_circuit(facts_probs.device), facts_probs)

    def update_semiring_weights(self, facts_probs):
        """weights[Term('p_ij')] = facts_probs[:, i-1, j]; new semiring per call (models/vael.py:227-243)."""
        for term in self.weights_dict:
            s = str(term)
            self.weights_dict[term] = facts_probs[:, int(s[-2]) - 1, int(s[-1])]
        self.semiring = GraphSemiring(facts_probs.shape[0], self.device)
        self.semiring.set_weights(self.weights_dict)

    def extract_query_probability(self, res):
        return res[next(reversed(res))][..., None]

    def extract_worlds_probability(self, res):
        """Stack digits(j,k) j-major, +1e-7, divide by the detached sum over the whole batch
        (models/vael.py:250-266)."""
        n = self.mlp.n_facts // 2
        if isinstance(res, EvaluationResult) and res.worlds is not None and res.worlds.shape[1] == n * n:
            p = res.worlds
        else:
            digits = Term("digits")
            p = torch.stack([res[digits(Constant(j), Constant(k))] for j in range(n) for k in range(n)], dim=1)
        p = p + 1e-7
        return p / p.detach().sum()


class MarioVAELModel(VAELModel):
    def __init__(self, encoder, decoder, mlp, latent_dims, model_dict, WQ_all, W_all, WQ_adm, W_adm, grid_dims,
                 dropout=None, is_train=True, device="cpu"):
        super().__init__(encoder=encoder, decoder=decoder, mlp=mlp, latent_dims=latent_dims, model_dict=model_dict,
                         dropout=dropout, is_train=is_train, device=device)
        self.grid_dims = grid_dims
        self.WQ_all, self.W_all, self.WQ_adm, self.W_adm = WQ_all, W_all, WQ_adm, W_adm
        self.eps = 1e-5
        for name, t, shape in (("W_all", W_all, (81, 18)), ("WQ_all", WQ_all, (81, 5)), ("W_adm", W_adm, (24, 18)),
                               ("WQ_adm", WQ_adm, (24, 4))):
            if t is not None:
                assert tuple(t.shape) == shape, (name, t.shape)
        self._psets = None

    def program_sets(self):
        """(world/query circuit in the probability semiring, admissible-world circuit in the log semiring)"""
        if self._psets is None:
            import numpy as np
            from .compiler import CompiledFamily
            from .mario_task import admissible_circuit, compile_mario_family, mario_tables
            fam = compile_mario_family(self.grid_dims)
            W_all, WQ_all, W_adm, WQ_adm = mario_tables(fam)
            for mine, given in ((W_all, self.W_all), (WQ_all, self.WQ_all), (W_adm, self.W_adm), (WQ_adm, self.WQ_adm)):
                if given is not None and not np.array_equal(mine, given.detach().cpu().numpy().astype(np.int8)):
                    raise RuntimeError("vael_b200: the Mario tables passed in differ from the compiled program's")
            adm = CompiledFamily(admissible_circuit(W_adm), {}, [], [], [])
            self._psets = (CompiledProgramSet(fam), CompiledProgramSet(adm))
        return self._psets

    def define_herbrand_base(self, n_facts):
        n = n_facts // 2
        eye = torch.eye(n)
        return torch.cat([eye.repeat_interleave(n, dim=0), eye.repeat(n, 1)], dim=1)

    def build_weights_dictionary(self, n_facts):
        return None

    def mario_facts(self, z_sym1, z_sym2=None):
        """softmax over the 9 positions, clip(1e-5, 0.9999) (models/vael.py:319-322): ONE kernel for both frames
        (facts_tile_kernel<2,9,clip>), returning the concatenated [B,18] fact matrix; a single frame -> [B,9]."""
        if z_sym2 is None:
            return facts_from_logits_clipped(self.mlp(z_sym1), 1, self.eps, 0.9999)
        logits = torch.cat([self.mlp(z_sym1), self.mlp(z_sym2)], dim=1)
        return facts_from_logits_clipped(logits, 2, self.eps, 0.9999)

    def forward(self, x, moves):
        x1, x2 = torch.split(x, x.shape[2] // 2, dim=2)
        mu1, logvar1 = self.encoder(x1)
        mu2, logvar2 = self.encoder(x2)
        z1, z2 = self.reparametrize(mu1, logvar1), self.reparametrize(mu2, logvar2)
        self.z_subsym1, self.z_subsym2 = z1[:, self.latent_dim_sym:], z2[:, self.latent_dim_sym:]
        if self.dropout and self.training:
            self.z_subsym1 = nn.functional.dropout(self.z_subsym1, p=self.dropout)
            self.z_subsym2 = nn.functional.dropout(self.z_subsym2, p=self.dropout)
        facts = self.mario_facts(z1[:, :self.latent_dim_sym], z2[:, :self.latent_dim_sym])
        self.facts_probs1, self.facts_probs2 = torch.split(facts, self.mlp.n_facts // 2, dim=1)

        self.worlds_probs = self.compute_worlds_prob_conj(facts)
        self.query_prob = self.compute_query_prob(moves, self.worlds_probs, facts=facts)
        self.adm_worlds_logprob = self.admissible_worlds_logprob(facts)
        world_h = self.sample_world(self.adm_worlds_logprob, self.W_adm.type(torch.float), is_log=True)
        world1, world2 = torch.split(world_h, self.mlp.n_facts // 2, dim=-1)
        x_recon1 = self.decode(self.z_subsym1, world1)
        x_recon2 = self.decode(self.z_subsym2, world2)
        return (x1.permute(0, 2, 3, 1), x_recon1.permute(0, 2, 3, 1), mu1, logvar1, x2.permute(0, 2, 3, 1),
                x_recon2.permute(0, 2, 3, 1), mu2, logvar2, self.query_prob)

    def compute_worlds_prob_conj(self, probs):
        """P(w) for the 81 worlds (models/vael.py:359-366: exp(log p @ W_all^T)) as the product
        level of the compiled circuit."""
        dc = self.program_sets()[0].device_circuit(probs.device)
        worlds, _ = circuit_forward(dc, probs, None)
        self._facts_of_worlds = probs
        return worlds

    def admissible_worlds_logprob(self, probs):
        """Unnormalised log-probability of the 24 admissible worlds (models/vael.py:368-375),
        log-semiring circuit over all 18 literals."""
        dc = self.program_sets()[1].device_circuit(probs.device)
        logp, _ = circuit_forward(dc, probs, None)
        return logp

    def compute_query_prob(self, labels, worlds_probs, facts=None):
        """P(move and constraint) / (sum_w P(w) + eps) (models/vael.py:377-399).  The per-sample
        python loop becomes a per-row query index into the circuit's query nodes."""
        move = torch.argmax(labels, dim=1).to(torch.int32)
        valid = (labels.sum(dim=1) == 1) & (labels.max(dim=1).values == 1)
        move = torch.where(valid, move, torch.full_like(move, -1))   # no exact (one-hot,1) row -> empty mask -> 0
        if facts is None:  # reference signature: the facts are the ones compute_worlds_prob_conj just saw
            facts = getattr(self, "_facts_of_worlds", None)
        if facts is None or facts.shape[0] != worlds_probs.shape[0]:
            raise RuntimeError("vael_b200: compute_query_prob needs the fact matrix the worlds were computed from")
        dc = self.program_sets()[0].device_circuit(facts.device)
        _, p_q = circuit_forward(dc, facts, move, want_worlds=False)
        return p_q[:, 0] / (torch.sum(worlds_probs, dim=1) + self.eps)

    def compute_query_prob_cond(self, q, worlds_probs):
        """P(move q | evidence) from conditioned admissible-world probabilities [.., 24] (models/vael.py:401-414)."""
        p_e = torch.sum(worlds_probs, dim=0) + self.eps
        return torch.sum(self.WQ_adm[:, int(q)] * worlds_probs, dim=0) / p_e

    def _evidence_circuit(self, device):
        if getattr(self, "_evid_pset", None) is None:
            from .compiler import CompiledFamily
            from .mario_task import admissible_evidence_circuit, mario_tables
            _, _, W_adm, WQ_adm = mario_tables(self.program_sets()[0].family)
            self._evid_pset = CompiledProgramSet(CompiledFamily(admissible_evidence_circuit(W_adm, WQ_adm), {}, [], [], []))
        return self._evid_pset.device_circuit(device)

    def compute_conditioned_worlds_prob(self, probs, evidence):
        """P(w | evidence move) over the 24 admissible worlds (models/vael.py:416-430), evaluation only:
        (p_w [w |= e] + eps) / (sum p_w [w |= e] + eps), the sum running over the whole tensor as in the reference.
        The masked products and P(e) come from the circuit kernel in evidence mode (VAEL_FLAG_EVIDENCE:
        P(w|e) = [w |= e] P(w) / P(e) per row, and P(e) as the query output)."""
        evidence = torch.as_tensor(evidence).to(probs.device).reshape(-1)
        if evidence.numel() != 4 or float(evidence.sum()) != 1.0 or float(evidence.max()) != 1.0:
            # no move matches an evidence row that is not one-hot: every world gets eps / eps
            return torch.ones(probs.shape[0], self.W_adm.shape[0], device=probs.device)
        move = int(torch.argmax(evidence))
        with torch.no_grad():
            cond, p_e = circuit_forward(self._evidence_circuit(probs.device), probs, move, evidence=True)
        return (cond * p_e + self.eps) / (p_e.sum() + self.eps)

    def encode_image(self, x):
        """Crisp position facts of one frame (models/vael.py:432-451), evaluation only."""
        with torch.no_grad():
            mu, logvar = self.encoder(x)
            z = self.reparametrize(mu, logvar)
            z_subsym = z[:, self.latent_dim_sym:]
            facts_probs_ = torch.softmax(self.mlp(z[:, :self.latent_dim_sym]), dim=1)
            facts_probs = torch.zeros_like(facts_probs_)
            mask = facts_probs_ >= facts_probs_.max()
            facts_probs[mask] = facts_probs_[mask]
            facts_probs[:, torch.where(facts_probs > 0.5)[1]] = 1.
        return z_subsym, facts_probs
