"""CPU tests of the host side: the circuit compiler against its bit-exact golden and against the
reference's own tables (w_q, Herbrand base, the four Mario .pt tables), the GraphSemiring mirror
against the reference op tables, and the C-ABI library (loads, exports every symbol the header
declares, rejects bad descriptions before touching CUDA).  No compute entry point is called."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
ARRAYS = ("node_kind", "node_arg", "level_ptr", "child_ptr", "child_idx", "const_val", "world_node", "query_node")


def load(name):
    return np.load(os.path.join(G, name + ".npz"), allow_pickle=False)


# ---------------------------------------------------------------------------
# compiler
# ---------------------------------------------------------------------------
def test_two_digit_circuit_arrays_bit_exact_against_golden(addition_family):
    g, c = load("circuit_addition_2digit"), addition_family.circuit
    for k in ARRAYS:
        a = getattr(c, k)
        assert a.dtype == g[k].dtype and np.array_equal(a, g[k]), k
    assert int(g["n_facts"]) == c.n_facts == 20 and int(g["norm_node"]) == c.norm_node
    assert list(g["fact_labels"]) == list(c.fact_labels)
    assert list(g["world_atoms"]) == list(c.world_atoms) and list(g["query_atoms"]) == list(c.query_atoms)
    assert str(load("graph_path")["circuit_sha256"]) == c.sha256()
    c.validate()


def test_world_indexing_and_query_table_match_reference(addition_family):
    """World w = d1*10+d2 (models/vael.py:218-221,255-258); the circuit's world->query membership is the
    reference's build_worlds_queries_matrix (mnist_task_VAEL.py:222-234); fact columns are p_1y then p_2y."""
    from vael_b200.mnist_task import build_worlds_queries_matrix
    c, g = addition_family.circuit, load("mnist_dense")
    assert list(c.world_atoms) == [f"digits({j},{k})" for j in range(10) for k in range(10)]
    assert list(c.query_atoms) == [f"addition(img,{s})" for s in range(19)]
    assert list(c.fact_labels) == [f"p_{p}{d}" for p in (1, 2) for d in range(10)]
    wq_ref = g["w_q"]
    assert np.array_equal(c.worlds_queries_matrix(), wq_ref[:, :19])
    assert np.array_equal(build_worlds_queries_matrix(2, 10).numpy(), wq_ref)
    # world node (j,k) is the product of the literals p_1j and p_2k
    for w in (0, 7, 42, 99):
        ch = c.children(int(c.world_node[w]))
        assert [int(c.node_arg[n]) for n in ch] == [w // 10, 10 + w % 10]
    # query members are listed in world order (the kernels add them in that order)
    pos = {int(n): w for w, n in enumerate(c.world_node)}
    for q in range(19):
        n = int(c.query_node[q])
        ws = [pos[int(x)] for x in c.children(n)] if int(c.node_kind[n]) == 4 else [pos[n]]
        assert ws == sorted(ws) and ws == [w for w in range(100) if w // 10 + w % 10 == q]


def test_program_family_views_and_text(addition_family):
    """38 programs (2 modes x 19 sums, mnist_task_VAEL.py:208-217) share one circuit; the text follows
    utils/mnist_utils/problog_model.py:27-54 (ADs, rules, digit query, then addition query | evidence)."""
    fam = addition_family
    assert len(fam.views) == 38
    v = fam.views[("query", 9)]
    assert v.mode == "query" and v.query_index == 9 and str(v.result_atoms[-1]) == "addition(img,9)"
    assert "query(digits(X,Y))." in v.text and "query(addition(img,9))." in v.text
    assert "addition(X,N) :- digit(X,1,N1), digit(X,2,N2), N is N1 + N2." in v.text
    e = fam.views[("evidence", 7)]
    assert e.mode == "evidence" and e.query_index == 7 and "evidence(addition(img,7))." in e.text
    assert len(e.result_atoms) == 100
    assert "0.1::digit" not in v.text and "p_10::digit(X,1,0)" in v.text and "p_29::digit(X,2,9)" in v.text


def test_herbrand_base_matches_reference():
    from vael_b200.networks import MNISTPairsMLP
    from vael_b200.vael import MNISTPairsVAELModel
    m = MNISTPairsVAELModel(None, None, MNISTPairsMLP(15, 20), (15, 8), None, None, device="cpu")
    assert np.array_equal(m.herbrand_base.numpy(), load("mnist_dense")["herbrand"])
    assert [str(t) for t in m.weights_dict] == [f"p_{p}{d}" for p in (1, 2) for d in range(10)]


def test_three_digit_family_structure():
    """The scaled program (BASELINE.json configs[4]): 3 ADs, 1000 worlds in mixed-radix order, sums 0..27."""
    from vael_b200.mnist_task import compile_addition_family
    c = compile_addition_family(3, 10).circuit
    assert (c.n_facts, c.n_worlds, c.n_queries) == (30, 1000, 28)
    assert list(c.world_atoms[:3]) == ["digits(0,0,0)", "digits(0,0,1)", "digits(0,0,2)"] and c.world_atoms[999] == "digits(9,9,9)"
    wq = c.worlds_queries_matrix()
    d = np.indices((10, 10, 10)).reshape(3, -1).sum(0)
    assert np.array_equal(wq, (d[:, None] == np.arange(28)[None]).astype(wq.dtype))
    c.validate()


def test_mario_tables_regenerated_from_the_rule_program_bit_exact():
    """utils/mario_utils/problog_matrices/3x3/{W_all,WQ_all,W_adm,WQ_adm}.pt regenerated by compiling the
    move-rule program (utils/mario_utils/problog_model.py:3-62) must equal the reference files bit for bit."""
    from vael_b200.mario_task import compile_mario_family, mario_tables
    g = load("mario")
    W_all, WQ_all, W_adm, WQ_adm = mario_tables(compile_mario_family((3, 3)))
    for mine, name in ((W_all, "W_all"), (WQ_all, "WQ_all"), (W_adm, "W_adm"), (WQ_adm, "WQ_adm")):
        assert mine.shape == g[name].shape and np.array_equal(mine.astype(np.int8), g[name]), name


def test_circuit_serialisation_round_trip(tmp_path, addition_family):
    from vael_b200.circuit import Circuit
    c = addition_family.circuit
    p = tmp_path / "c.npz"
    c.save(p)
    c2 = Circuit.load(p)
    assert c2.sha256() == c.sha256() and c2.fnv1a64() == c.fnv1a64()
    bad = Circuit.load(p)
    bad.child_idx = bad.child_idx.copy()
    bad.child_idx[-1] = bad.n_nodes - 1          # a child in its own level
    with pytest.raises(ValueError):
        bad.validate()


def test_compiler_rejects_what_it_cannot_compile():
    from vael_b200.compiler import CompileError, compile_family
    with pytest.raises(CompileError):
        compile_family({0: "a :- b.\nquery(a)."})                       # no probabilistic facts
    with pytest.raises(CompileError):
        compile_family({0: "p_1::a; p_2::b.\nevidence(a, false).\nquery(b)."})   # negative evidence


# ---------------------------------------------------------------------------
# GraphSemiring mirror (utils/graph_semiring.py:5-60) against the reference op tables
# ---------------------------------------------------------------------------
def test_graph_semiring_mirror_matches_reference_tables():
    from vael_b200.graph_semiring import GraphSemiring
    g = load("semiring_ops")
    T = lambda k: torch.from_numpy(g[k])
    s = GraphSemiring(batch_size=8, device=torch.device("cpu"))
    assert s.eps == 1e-12 and s.batch_size == 8
    cases = {"pp": ("a", "b_plain"), "pz": ("a", "b_zero"), "zp": ("a_zero", "b_plain"), "po": ("a", "b_one"),
             "op": ("a_one", "b_plain")}
    for tag, (x, y) in cases.items():
        if f"plus_{tag}" in g.files:
            assert torch.equal(s.plus(T(x), T(y)), T(f"plus_{tag}")), tag
            assert torch.equal(s.times(T(x), T(y)), T(f"times_{tag}")), tag
    assert torch.equal(s.negate(T("a")), T("negate_a"))
    assert torch.equal(s.one(), T("one")) and torch.equal(s.zero(), T("zero"))
    s.set_weights({"k": T("a")})
    assert torch.equal(s.value("k"), T("value_hit")) and s.value(0.25) == 0.25
    # the elementwise variant is what the kernels implement: no batch-wide shortcut
    e = GraphSemiring(8, torch.device("cpu"), elementwise=True)
    assert torch.equal(e.plus(T("a"), T("b_zero")), T("a") + T("b_zero"))


def test_fact_matrix_recovers_the_contiguous_block_without_a_copy():
    from vael_b200.graph_semiring import GraphSemiring
    from vael_b200.logic import Term
    facts = torch.rand(6, 2, 10, requires_grad=True)
    fp = facts * 1.0
    terms = [Term(f"p_{p}{d}") for p in (1, 2) for d in range(10)]
    s = GraphSemiring(6)
    s.set_weights({t: fp[:, i // 10, i % 10] for i, t in enumerate(terms)})   # models/vael.py:236-240
    m = s.fact_matrix(terms)
    assert m.shape == (6, 20) and m.data_ptr() == fp.data_ptr() and m.requires_grad
    s.set_weights({t: torch.rand(6) for t in terms})
    assert s.fact_matrix(terms).shape == (6, 20)
    with pytest.raises(KeyError):
        GraphSemiring(6).fact_matrix(terms)


# ---------------------------------------------------------------------------
# C-ABI library
# ---------------------------------------------------------------------------
def header_symbols():
    text = open(os.path.join(ROOT, "include", "vael_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"VAEL_API\s+[\w\s\*]+?\b(vael_\w+)\s*\(", text)


def test_library_loads_and_exports_every_declared_symbol():
    from vael_b200 import _lib
    names = header_symbols()
    assert len(names) >= 19 and len(set(names)) == len(names)
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"libvael_b200.so does not export {n}"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), "ctypes prototypes and include/vael_b200.h disagree"
    assert lib.vael_abi_version() == _lib.ABI_VERSION == 2
    assert lib.vael_elbo_workspace_bytes() > 0 and lib.vael_launch_count() >= 0


def test_library_validates_descriptions_before_touching_cuda(addition_family):
    from vael_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.vael_circuit_create(None, C.byref(h)) == 1 and b"null" in lib.vael_last_error()
    c = addition_family.circuit
    import copy
    bad = copy.copy(c)
    bad.child_idx = c.child_idx.copy()
    bad.child_idx[-1] = c.n_nodes - 1
    assert lib.vael_circuit_create(C.byref(_lib.make_desc(bad)), C.byref(h)) == 1
    assert b"strictly lower level" in lib.vael_last_error()
    bad2 = copy.copy(c)
    bad2.world_node = c.world_node.copy()
    bad2.world_node[3] = 10 ** 6
    assert lib.vael_circuit_create(C.byref(_lib.make_desc(bad2)), C.byref(h)) == 1
    assert lib.vael_circuit_fast_kind(None) == 0 and lib.vael_circuit_hash(None) == 0
    lib.vael_circuit_destroy(None)


def test_product_path_has_no_cpu_fallback(addition_family):
    """CPU tensors are refused loudly; nothing under vael_b200/ imports the oracle."""
    from vael_b200.ops import DeviceCircuit, facts_from_logits
    with pytest.raises(RuntimeError, match="CPU|CUDA|cuda"):
        facts_from_logits(torch.randn(4, 20), 2)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CUDA device"):
            DeviceCircuit(addition_family.circuit)
    pkg = os.path.join(ROOT, "vael_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


# ---------------------------------------------------------------------------
# encoder / decoder / MLP mirrors: checkpoint layout and forward semantics of the reference modules
# ---------------------------------------------------------------------------
def _mirror(name, **kw):
    from vael_b200 import networks
    return getattr(networks, name)(**kw)


def test_network_state_dict_layouts_match_reference_checkpoints():
    """Names and shapes of every parameter at the shipped widths (what {'model': state_dict} checkpoints
    of the reference hold, utils/early_stopping.py:53-65)."""
    import json
    lay = json.load(open(os.path.join(G, "network_layouts.json")))
    mine = {
        "MNISTPairsEncoder": _mirror("MNISTPairsEncoder", hidden_channels=64, latent_dim=23),
        "MNISTPairsDecoder": _mirror("MNISTPairsDecoder", label_dim=20, hidden_channels=64, latent_dim=8),
        "MNISTPairsMLP": _mirror("MNISTPairsMLP", in_features=15, n_facts=20),
        "MarioEncoder": _mirror("MarioEncoder", img_channels=3, hidden_channels=64, latent_dim=48),
        "MarioDecoder": _mirror("MarioDecoder", img_channels=3, hidden_channels=64, latent_dim_sub=30, label_dim=9),
        "MarioMLP": _mirror("MarioMLP", in_features=18, n_facts=18, hidden_channels=32),
    }
    for name, m in mine.items():
        got = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert got == lay[name], name
    assert sum(p.numel() for n in ("MNISTPairsEncoder", "MNISTPairsDecoder", "MNISTPairsMLP")
               for p in mine[n].parameters()) == 1848467
    assert sum(p.numel() for n in ("MarioEncoder", "MarioDecoder", "MarioMLP") for p in mine[n].parameters()) == 10354412


def test_networks_load_reference_weights_and_reproduce_the_reference_forward():
    g = np.load(os.path.join(G, "networks.npz"))
    ctor = {
        "MNISTPairsEncoder": dict(hidden_channels=4, latent_dim=23),
        "MNISTPairsDecoder": dict(label_dim=20, hidden_channels=4, latent_dim=8),
        "MNISTPairsMLP": dict(in_features=15, n_facts=20),
        "MarioEncoder": dict(img_channels=3, hidden_channels=2, latent_dim=48),
        "MarioDecoder": dict(img_channels=3, hidden_channels=2, latent_dim_sub=30, label_dim=9),
        "MarioMLP": dict(in_features=18, n_facts=18, hidden_channels=32),
    }
    for name, kw in ctor.items():
        m = _mirror(name, **kw)
        sd = {k[len(name) + 3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(name + "/w/")}
        m.load_state_dict(sd, strict=True)
        m.eval()
        with torch.no_grad():
            y = m(torch.from_numpy(g[name + "/x"]))
        ys = y if isinstance(y, (tuple, list)) else (y,)
        for i, t in enumerate(ys):
            np.testing.assert_allclose(t.numpy(), g[f"{name}/y{i}"], rtol=1e-5, atol=1e-6, err_msg=name)


# ---------------------------------------------------------------------------
# import shim: the reference's own compile calls end in this repo's compiler
# ---------------------------------------------------------------------------
def test_problog_shim_compiles_the_reference_program_text():
    """SDD.create_from(LogicDAG.create_from(LogicFormula.create_from(text))) as
    utils/mnist_utils/mnist_task_VAEL.py:214-216 writes it, on the text this repo's generator emits (the
    next test checks that text against the reference's own generator where the reference tree exists)."""
    from vael_b200 import compat
    from vael_b200.mnist_task import addition_rules, create_facts, define_ProbLog_model
    compat.install(force=True)
    from problog.formula import LogicDAG, LogicFormula
    from problog.logic import Constant, Term
    from problog.program import PrologString
    from problog.sdd_formula import SDD
    facts = create_facts(2, n_digits=10)
    for mode, add in (("query", 9), ("evidence", 7)):
        text = define_ProbLog_model(facts, addition_rules(2), label=add, digit_query="digits(X,Y)", mode=mode)
        sdd = SDD.create_from(LogicDAG.create_from(LogicFormula.create_from(PrologString(text))))
        assert sdd.mode == mode and hasattr(sdd, "evaluate")
        circ = sdd.pset.circuit
        assert circ.n_worlds == 100 and circ.n_queries == 1 and circ.n_facts == 20
        assert [str(a) for a in sdd.view.result_atoms[:100]] == [f"digits({j},{k})" for j in range(10) for k in range(10)]
        assert sdd.view.result_atoms[0] == Term("digits")(Constant(0), Constant(0))
        members = circ.worlds_queries_matrix()[:, 0]
        assert [w for w in range(100) if members[w]] == [w for w in range(100) if w // 10 + w % 10 == add]
    sdd2 = SDD.create_from(LogicFormula.create_from(text))          # a bare string works too
    assert sdd2.pset.circuit.sha256() == circ.sha256()


@pytest.mark.skipif(not os.path.exists("/root/reference/utils/mnist_utils/problog_model.py"),
                    reason="the reference tree exists in the build container only")
def test_program_text_is_what_the_reference_generator_writes():
    """utils/mnist_utils/problog_model.py imported UNMODIFIED through the shim: same facts, same program text."""
    import importlib.util
    from vael_b200 import compat, mnist_task
    compat.install(force=True)
    spec = importlib.util.spec_from_file_location("ref_problog_model", "/root/reference/utils/mnist_utils/problog_model.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    assert ref.create_facts(2, n_digits=10) == mnist_task.create_facts(2, n_digits=10)
    rules = "addition(X,N) :- digit(X,1,N1), digit(X,2,N2), N is N1 + N2.\ndigits(X,Y):-digit(img,1,X), digit(img,2,Y)."
    assert rules == mnist_task.addition_rules(2)                     # mnist_task_VAEL.py:206
    facts = ref.create_facts(2, n_digits=10)
    for mode in ("query", "evidence"):
        for add in (0, 9, 18):
            assert ref.define_ProbLog_model(facts, rules, label=add, digit_query="digits(X,Y)", mode=mode) == \
                mnist_task.define_ProbLog_model(facts, rules, label=add, digit_query="digits(X,Y)", mode=mode)


def _small_models():
    import torch
    from vael_b200 import networks as N
    from vael_b200.mario_task import compile_mario_family, mario_tables
    from vael_b200.vael import MarioVAELModel, MNISTPairsVAELModel
    dev = torch.device("cpu")
    mnist = MNISTPairsVAELModel(N.MNISTPairsEncoder(hidden_channels=4, latent_dim=23),
                                N.MNISTPairsDecoder(label_dim=20, hidden_channels=4, latent_dim=8),
                                N.MNISTPairsMLP(in_features=15, n_facts=20), (15, 8), None, torch.zeros(100, 20), dropout=0.5,
                                is_train=True, device=dev)
    W_all, WQ_all, W_adm, WQ_adm = (torch.from_numpy(t.astype("float32")) for t in mario_tables(compile_mario_family((3, 3))))
    mario = MarioVAELModel(N.MarioEncoder(img_channels=3, hidden_channels=2, latent_dim=48),
                           N.MarioDecoder(img_channels=3, hidden_channels=2, latent_dim_sub=30, label_dim=9),
                           N.MarioMLP(in_features=18, n_facts=18, hidden_channels=32), (18, 30), None, WQ_all, W_all, WQ_adm,
                           W_adm, (3, 3), dropout=0.5, is_train=True, device=dev)
    return {"MNISTPairsVAELModel": (mnist, "ref_checkpoint_mnist.pt"), "MarioVAELModel": (mario, "ref_checkpoint_mario.pt")}


def test_reference_checkpoints_load_and_round_trip(tmp_path):
    """{'model', 'optimizer'} checkpoints written by the reference's own classes (utils/early_stopping.py:53-65;
    tests/golden/make_checkpoint_golden.py) load strictly into this repo's models with their Adam state, and a
    checkpoint written here has exactly the reference's keys and shapes."""
    import json
    import torch
    from vael_b200.checkpoint import load_checkpoint, save_checkpoint
    lay = json.load(open(os.path.join(G, "checkpoint_layouts.json")))
    for name, (model, fname) in _small_models().items():
        assert {k: list(v.shape) for k, v in model.state_dict().items()} == lay[name], name
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        ckpt = load_checkpoint(os.path.join(G, fname), model, opt, map_location="cpu")
        assert set(ckpt) == {"model", "optimizer"}
        for k, v in model.state_dict().items():
            assert torch.equal(v, ckpt["model"][k]), k
        st = opt.state_dict()["state"]
        assert len(st) == len(list(model.parameters())) and all(int(s["step"]) == 1 for s in st.values())
        out = save_checkpoint(model, opt, str(tmp_path / name / "checkpoint.pt"))
        again = torch.load(out, weights_only=True)
        assert list(again) == ["model", "optimizer"]
        assert {k: list(v.shape) for k, v in again["model"].items()} == lay[name]
        assert again["optimizer"]["param_groups"][0]["lr"] == ckpt["optimizer"]["param_groups"][0]["lr"]
    with pytest.raises(ValueError):
        torch.save({"weights": {}}, str(tmp_path / "bad.pt"))
        load_checkpoint(str(tmp_path / "bad.pt"), model)


def test_learning_curve_files_round_trip(tmp_path):
    """train_info.npy / validation_info.npy: dict of per-epoch means under the reference's keys
    (utils/mnist_utils/train.py:74-93,160-164,239-240), read back the way its analysis code reads them."""
    from vael_b200.checkpoint import CURVE_KEYS, append_epoch, load_curves, save_curves
    tr, va = {}, {}
    for e in range(3):
        append_epoch(tr, {k: 10.0 - e + i for i, k in enumerate(CURVE_KEYS)})
        append_epoch(va, {k: 11.0 - e + i for i, k in enumerate(CURVE_KEYS)})
    save_curves(str(tmp_path), tr, va)
    tr2, va2 = load_curves(str(tmp_path))
    assert list(tr2) == list(CURVE_KEYS) and tr2 == tr and va2 == va and len(tr2["elbo"]) == 3


def test_as_channels_last_gives_one_channel_images_explicit_nhwc_strides():
    """For C = 1 torch keeps NCHW strides under .contiguous(channels_last) and the convolution heuristics then see
    an NCHW tensor; train.as_channels_last hands out the same bytes with strides (H*W, 1, W, 1)."""
    from vael_b200.train import as_channels_last
    x = torch.arange(2 * 1 * 3 * 5, dtype=torch.float32).reshape(2, 1, 3, 5)
    plain = x.contiguous(memory_format=torch.channels_last)
    assert plain.stride() == (15, 15, 5, 1)                       # ambiguous: what torch leaves us with
    y = as_channels_last(x)
    assert y.stride() == (15, 1, 5, 1) and torch.equal(y, x) and y.data_ptr() == x.data_ptr()
    assert y.is_contiguous(memory_format=torch.channels_last)
    assert y.suggest_memory_format() == torch.channels_last if hasattr(y, "suggest_memory_format") else True
    z = as_channels_last(torch.randn(2, 3, 4, 4))                 # C > 1: the ordinary conversion
    assert z.stride() == (48, 1, 12, 3)


def test_reference_copy_under_baseline_ref_is_verbatim():
    """scripts/install_ref.py copies EleMisi/VAEL verbatim to baseline/_ref (git-ignored, ships with gpurun); the
    sha256 manifest of that copy is committed.  Skipped where the copy has not been made."""
    from tests import refimport
    if not refimport.available():
        pytest.skip("baseline/_ref not installed (python scripts/install_ref.py)")
    assert refimport.modified_files() == []


def test_loss_functions_refuse_cpu_tensors():
    """No silent CPU path: the ELBO terms come from the kernel or not at all."""
    import torch
    from vael_b200.train import loss_function, mario_loss_function
    x = torch.zeros(2, 1, 4, 4)
    mu = torch.zeros(2, 3)
    for rl in ("LAPLACE", "MSE", "BCE"):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            loss_function(x, x, mu, mu, torch.ones(2, 1), rec_loss=rl)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mario_loss_function(x, x, mu, mu, x, x, mu, mu, torch.ones(2), rec_loss="LAPLACE")


def test_mnist_model_checks_the_worlds_queries_matrix():
    """A w_q that is not the compiled program's table raises instead of being ignored."""
    import torch
    from vael_b200.mnist_task import build_worlds_queries_matrix
    from vael_b200.networks import MNISTPairsMLP
    from vael_b200.vael import MNISTPairsVAELModel
    w_q = build_worlds_queries_matrix(2, 10)
    ok = MNISTPairsVAELModel(None, None, MNISTPairsMLP(in_features=15, n_facts=20), (15, 8), None, w_q, device="cpu")
    assert ok.program_set().circuit.n_queries == 19 and ok._n_wq_cols == 20
    bad = w_q.clone()
    bad[[3, 4]] = bad[[4, 3]]
    m = MNISTPairsVAELModel(None, None, MNISTPairsMLP(in_features=15, n_facts=20), (15, 8), None, bad, device="cpu")
    with pytest.raises(RuntimeError, match="worlds-queries matrix"):
        m.program_set()


def test_generated_circuit_kernels_compile_for_sm100a(addition_family):
    """circuit_jit.cu: the straight-line source generated for a circuit compiles with NVRTC for sm_100a (no GPU
    needed: the module is not loaded); circuits outside the generator's scope are reported as such.  The 3-digit
    circuit takes the column-chunked variant (2-D tensor-map box copies, 9 k lines of source, ~20 s of NVRTC)."""
    import ctypes
    from vael_b200 import _lib
    from vael_b200.mario_task import compile_mario_family
    from vael_b200.mnist_task import compile_addition_family
    lib = _lib.load()
    buf = ctypes.create_string_buffer(1 << 16)
    from vael_b200.mario_task import admissible_circuit, mario_tables
    adm = admissible_circuit(mario_tables(compile_mario_family((3, 3)))[2])        # log semiring: out of scope
    wide = compile_addition_family(3, 10).circuit
    for circ, want in ((addition_family.circuit, 0), (compile_mario_family((3, 3)).circuit, 0), (adm, 1), (wide, 0)):
        d = _lib.make_desc(circ)
        rc = lib.vael_jit_compile_check(ctypes.byref(d), buf, len(buf))
        if rc == 2:
            pytest.skip("NVRTC not available: " + buf.value.decode())
        assert rc == want, buf.value.decode()[:2000]


def test_device_pairs_loader_serves_the_reference_batch_composition():
    """vael_b200/data.py: one world per batch, worlds without replacement, every sample once per epoch, labels
    [d1, d2, sum, -1], whole-array standardisation (utils/mnist_utils/mnist_addition_dataset.py:37,67-121)."""
    import torch
    from vael_b200.data import SyntheticPairsLoader
    ld = SyntheticPairsLoader(samples_per_world=12, batch_size=4, digits=3, device="cpu", seed=7)
    assert len(ld.worlds) == 9 and len(ld) == 27 and ld.samples_x_world == 12
    seen = []
    for images, labels in ld:
        assert images.shape == (4, 28, 56) and images.dtype == torch.float32 and labels.shape == (4, 4)
        assert (labels == labels[0]).all() and int(labels[0, 2]) == int(labels[0, 0] + labels[0, 1]) and int(labels[0, 3]) == -1
        seen.append(tuple(labels[0, :2].tolist()))
    assert sorted(set(seen)) == sorted(ld.worlds) and all(seen.count(c) == 3 for c in ld.worlds)
    with pytest.raises(IndexError):
        ld[0]
    ld.reset_counter()
    images, _ = ld[0]
    raw = ld.images.float()
    assert abs(float(raw.mean()) - ld.mean) < 1e-4 and abs(((images * ld.std + ld.mean) - images * ld.std - ld.mean).abs().max()) < 1e-3
