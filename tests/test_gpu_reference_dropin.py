"""The drop-in proof with the reference's OWN classes on the GPU.

baseline/_ref holds EleMisi/VAEL verbatim (scripts/install_ref.py; sha256 manifest committed under
tests/golden/).  These tests import the reference's MNISTPairsVAELModel / MarioVAELModel, its unmodified
utils/graph_semiring.py:GraphSemiring, its build_model_dict (utils/mnist_utils/mnist_task_VAEL.py:203-219: the
LogicFormula -> LogicDAG -> SDD.create_from chain) and its evaluation code (utils/mnist_utils/metrics.py), put
`vael_b200.compat` where `problog` would be, and run them on CUDA: every `sdd.evaluate(semiring=...)` is then
served by the circuit kernels of libvael_b200.so.  Results are held to the goldens the reference itself wrote
(tests/golden/graph_path.npz) at the north-star tolerance 1e-5.
"""
import os

import numpy as np
import pytest
import torch

from tests import refimport

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref(cuda_device):
    if not refimport.available():
        pytest.fail("baseline/_ref is missing: run `python scripts/install_ref.py` in the build container "
                    "(it ships to the GPU box with the gpurun snapshot)")
    refimport.install()
    import models.vael as vael
    import models.vael_networks as nets
    import utils.graph_semiring as gs
    import utils.mnist_utils.metrics as metrics
    import utils.mnist_utils.mnist_task_VAEL as task
    import utils.mnist_utils.train as train
    for m in (vael, nets, gs, metrics, task, train):
        assert m.__file__.startswith(refimport.REF_DIR), m.__file__
    return dict(vael=vael, nets=nets, gs=gs, metrics=metrics, task=task, train=train)


@pytest.fixture(scope="module")
def ref_model(ref, cuda_device):
    """Built exactly as run_mnist_vael does (mnist_task_VAEL.py:260-262, :292-301)."""
    task, nets, vael = ref["task"], ref["nets"], ref["vael"]
    model_dict = task.build_model_dict(2, 10)                    # the reference's own three-call chain
    w_q = task.build_worlds_queries_matrix(2, 10).to(cuda_device)
    torch.manual_seed(0)
    model = vael.MNISTPairsVAELModel(encoder=nets.MNISTPairsEncoder(hidden_channels=64, latent_dim=23, dropout=0.5),
                                     decoder=nets.MNISTPairsDecoder(label_dim=20, hidden_channels=64, latent_dim=8, dropout=0.5),
                                     mlp=nets.MNISTPairsMLP(in_features=15, n_facts=20), latent_dims=(15, 8),
                                     model_dict=model_dict, w_q=w_q, dropout=0.5, is_train=True, device=cuda_device)
    return model.to(cuda_device)


def test_reference_copy_is_unmodified():
    assert refimport.available()
    assert refimport.modified_files() == []


def test_reference_model_evidence_and_query_inference_on_cuda(ref, ref_model, cuda_device):
    """models/vael.py:119-135 (problog_inference_with_evidence) and :98-117 (the generic problog_inference the
    MNIST subclass shadows) executed by the reference's code with the reference's GraphSemiring; the compiled
    programs in model_dict answer `evaluate` from the CUDA kernels."""
    from vael_b200 import _lib
    g = np.load(os.path.join(GOLD, "graph_path.npz"))
    facts = torch.from_numpy(g["facts"]).to(cuda_device)                          # [32,2,10]
    n0 = _lib.launch_count()
    worlds_e = ref_model.problog_inference_with_evidence(facts, 7)
    assert _lib.launch_count() > n0, "no vael_b200 kernel ran"
    assert type(ref_model.semiring) is ref["gs"].GraphSemiring                     # bound by the reference (:242)
    assert worlds_e.is_cuda and worlds_e.shape == (32, 100)
    np.testing.assert_allclose(worlds_e.cpu().numpy(), g["evidence_extract_worlds"], rtol=1e-5, atol=1e-12)
    # the generic query-mode path of the base class, called unbound so the dense override does not shadow it
    qp, worlds = ref["vael"].VAELModel.problog_inference(ref_model, facts, 9, None)
    np.testing.assert_allclose(qp.cpu().numpy(), g["extract_query"], rtol=1e-5)
    np.testing.assert_allclose(worlds.cpu().numpy(), g["extract_worlds"], rtol=1e-5, atol=1e-12)
    # ... and it agrees with the reference's own dense restatement (models/vael.py:208-225) on the same facts
    qd, wd = ref_model.problog_inference(facts, labels=torch.tensor([[0, 0, 9, -1]] * 32))
    np.testing.assert_allclose(qp.cpu().numpy(), qd.cpu().numpy(), rtol=1e-5)
    # extract_worlds_probability: (P(w) + 1e-7) / sum over the WHOLE batch (models/vael.py:260-266)
    np.testing.assert_allclose(worlds.cpu().numpy(), ((wd + 1e-7) / (wd + 1e-7).sum()).cpu().numpy(), rtol=2e-5)


def test_reference_model_gradients_flow_through_the_compiled_program(ref, ref_model, cuda_device):
    """autograd of the reference's generic path = the adjoint kernel (utils/mnist_utils/train.py:147); compared at
    the logits (SURVEY.md hard part 3) with autograd through the reference's dense form."""
    torch.manual_seed(3)
    z = torch.randn(32, 15, device=cuda_device)
    gq = torch.randn(32, 1, device=cuda_device)

    def grads(path):
        for p in ref_model.mlp.parameters():
            p.grad = None
        zz = z.clone().requires_grad_(True)
        facts = ref_model.compute_facts_probability(zz)
        qp, _ = path(facts)
        (qp * gq).sum().backward()
        return zz.grad.clone()

    g_graph = grads(lambda f: ref["vael"].VAELModel.problog_inference(ref_model, f, 11, None))
    g_dense = grads(lambda f: ref_model.problog_inference(f, query=11))
    assert float((g_graph - g_dense).abs().max()) <= 1e-4 * float(g_dense.abs().max())


def test_reference_generative_ability_runs_on_the_kernels(ref, ref_model, cuda_device):
    """utils/mnist_utils/metrics.py:88-132, the one live caller of the SDD path in the reference: 19 evidence
    programs x (facts -> evaluate -> gumbel -> herbrand -> decode -> classify)."""
    from vael_b200 import _lib

    class Clf(torch.nn.Module):                      # stands in for MNIST_classifier.pt: log-probabilities [1,10]
        def forward(self, x):
            return torch.log_softmax(x[:, :10], dim=1)

    n0 = _lib.launch_count()
    acc = ref["metrics"].generative_ability(ref_model, Clf().to(cuda_device), 16)
    assert 0.0 <= acc <= 1.0
    assert _lib.launch_count() - n0 >= 19            # one circuit evaluation per evidence at least
    assert type(ref_model.semiring) is ref["gs"].GraphSemiring


def test_reference_training_loop_body_drives_this_repos_model(ref, cuda_device):
    """The other direction: the reference's loss_function (utils/mnist_utils/train.py:13-51) and batch-loop body
    (:105-150) around THIS repo's MNISTPairsVAELModel built from the reference's networks and model_dict."""
    from vael_b200.vael import MNISTPairsVAELModel
    task, nets, rtrain = ref["task"], ref["nets"], ref["train"]
    torch.manual_seed(1)
    model = MNISTPairsVAELModel(nets.MNISTPairsEncoder(hidden_channels=64, latent_dim=23, dropout=0.5),
                                nets.MNISTPairsDecoder(label_dim=20, hidden_channels=64, latent_dim=8, dropout=0.5),
                                nets.MNISTPairsMLP(in_features=15, n_facts=20), (15, 8), task.build_model_dict(2, 10),
                                task.build_worlds_queries_matrix(2, 10).to(cuda_device), dropout=0.5, is_train=True,
                                device=cuda_device).to(cuda_device)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    model.semiring = ref["gs"].GraphSemiring(batch_size=30, device=model.device)    # train.py:101
    x = torch.randn(30, 1, 28, 56, device=cuda_device)
    labels = torch.tensor([[3, 4, 7, -1]] * 30, device=cuda_device)
    losses = []
    for _ in range(5):
        opt.zero_grad()
        recon, mu, logvar, add_prob = model(x, labels)
        loss, *_ = rtrain.loss_function(recon, x, mu, logvar, add_prob, model=model, labels=labels, query=True,
                                        recon_w=1e-1, kl_w=1e-5, query_w=1.0, sup_w=0.0, sup=False, rec_loss="LAPLACE")
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
