"""Drop-in check: the REFERENCE's own `JackRenderer` / `JackRendererBenchmark` objects
(`ReTiSAR/_jack_renderer.py`, imported unmodified from the staged copy `oracle/_ref`, JACK itself
replaced by `oracle/fake_jack.py`) are driven once with the reference's convolver classes and once
with this repository's classes swapped in at the import sites INTEGRATION.md names
(`_jack_renderer.py:3-5`).  Both must deliver the same blocks (bar: 1e-5 of full scale)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5

CLIENT_KW = dict(is_single_precision=True, ir_trunc_db=None, is_detect_clipping=False,
                 is_measure_levels=False, is_measure_load=False, is_main_client=False,
                 is_disable_file_logger=True, is_disable_logger=True)


def _reference():
    from oracle import ref_harness as rh

    if not rh.reference_available():
        pytest.skip("no reference package staged in oracle/_ref (run __graft_entry__.build() where "
                    "/root/reference exists)")
    rh.import_reference()
    import ReTiSAR._jack_renderer as jr

    return rh, jr


def _alive(client):
    """The client process is never started here (no JACK server): answer `is_alive()` so that the
    control methods (`_check_alive`, `_jack_client.py:897-915`) act on the object in this process."""
    client.is_alive = lambda: True
    return client


def _swap_in_repo_classes(monkeypatch, jr):
    """The two-line change of INTEGRATION.md, applied to the imported reference module."""
    import retisar_b200 as ours
    from retisar_b200 import _convolver as oc, _filter_set as of

    monkeypatch.setattr(jr, "Convolver", ours.Convolver)
    monkeypatch.setattr(jr, "FilterSet", ours.FilterSet)
    for name in ("AdjustableFdConvolver", "AdjustableShConvolver"):
        monkeypatch.setattr(jr, name, getattr(oc, name))
    for name in ("FilterSetMiro", "FilterSetSofa"):
        monkeypatch.setattr(jr, name, getattr(of, name))


def test_reference_benchmark_client_with_repo_convolvers(monkeypatch):
    """`JackRendererBenchmark.__init__` / `add_convolver` / `_process` (`_jack_renderer.py:467-531`)."""
    rh, jr = _reference()
    from retisar_b200 import _convolver as oc

    kw = dict(name="bench", block_length=256, filter_length=1024, **CLIENT_KW)
    r_ref = _alive(jr.JackRendererBenchmark(**kw))
    _swap_in_repo_classes(monkeypatch, jr)
    r_new = _alive(jr.JackRendererBenchmark(**kw))
    assert type(r_new._convolver) is oc.OverlapSaveConvolver
    assert type(r_ref._convolver).__module__.startswith("ReTiSAR")
    for r in (r_ref, r_new):
        for _ in range(3):
            r.add_convolver()            # copy(self._convolver): independent clones
        r.set_client_passthrough(False)  # as `_run_benchmark._create_client` does (:440-444)
        assert len(r._convolvers) == 4 and len({id(c) for c in r._convolvers}) == 4
    rng = np.random.default_rng(1)
    for t in range(6):
        x = (rng.standard_normal((2, 256)) * 0.1).astype(np.float32)
        y_ref, y_new = r_ref._process(x.copy()), r_new._process(x.copy())
        assert y_new.shape == y_ref.shape == (2, 256) and y_new.dtype == np.float32
        assert np.abs(y_new - y_ref).max() <= TOL
        y_new *= 0.5  # the JACK client scales the returned block in place (`_jack_client.py:795`)


def test_reference_renderer_client_fir_filter_and_controls(monkeypatch):
    """`JackRenderer` with an ndarray FIR set (`_init_convolver`, `:88-164`), `_process` (`:296-315`),
    `set_client_passthrough` (`:317-349`) and the pause protocol `filter_block(None)`."""
    rh, jr = _reference()
    from retisar_b200 import synthetic as syn

    irs = syn.decaying_noise_irs((2, 1500), 5, scale=0.2)
    kw = dict(name="fir", block_length=128, filter_name=irs, filter_type="FIR_MULTICHANNEL", **CLIENT_KW)
    import oracle.fake_jack as fj

    monkeypatch.setattr(fj, "DEFAULT_BLOCKSIZE", 128)
    r_ref = _alive(jr.JackRenderer(**kw))
    _swap_in_repo_classes(monkeypatch, jr)
    r_new = _alive(jr.JackRenderer(**kw))
    assert r_new._convolver._plan.P == r_ref._convolver._filter.get_filter_blocks_fd().shape[0] == 12
    rng = np.random.default_rng(2)
    for t in range(30):
        if t == 14:
            assert r_ref.set_client_passthrough(True) is True and r_new.set_client_passthrough(True) is True
        if t == 17:
            r_ref.set_client_passthrough(False); r_new.set_client_passthrough(False)
        if t == 22:  # paused pipeline: None in, None out (`_jack_client.py:730-731`)
            assert r_new._process(None) is None and r_ref._process(None) is None
        x = (rng.standard_normal((2, 128)) * 0.1).astype(np.float32)
        y_ref, y_new = r_ref._process(x.copy()), r_new._process(x.copy())
        assert np.abs(y_new - y_ref).max() <= TOL, t


def test_reference_renderer_client_sh_renderer(monkeypatch):
    """The binaural renderer behind the reference client: `JackRenderer._process`,
    `set_client_crossfade` (`:351-377`) and the tracker array read per block; the convolvers are
    built from the same synthetic em32 / HRIR tables on both sides.  The reference client runs
    first, completely, with the reference's classes; then the module is patched and the same
    session is replayed on this repository's classes."""
    rh, jr = _reference()
    import ReTiSAR
    from retisar_b200 import Convolver
    from retisar_b200 import workloads as wl

    tb = wl.sh_tables("c2", hrir_dirs=240)
    B, order, C = tb["B"], tb["order"], tb["C"]
    irs = np.zeros((2, B), np.float32)
    irs[:, 0] = 1
    kw = dict(name="sh", block_length=B, filter_name=irs, filter_type="FIR_MULTICHANNEL", **CLIENT_KW)
    tracker = rh.make_tracker()
    azim = ReTiSAR.HeadTracker.DataIndex.AZIM
    T = 12
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((T, C, B)) * 0.1).astype(np.float32)

    def session(client):
        client._is_passthrough = False
        outs = []
        for t in range(T):
            if t == 5:
                assert client.set_client_crossfade(False) is False
            if t == 8:
                assert client.set_client_crossfade(True) is True
            tracker[azim] = float(37.5 * t - 100.0)
            outs.append(client._process(x[t].copy()).copy())
        return np.stack(outs)

    r_ref = _alive(jr.JackRenderer(**kw))
    hr, ar = tb["hrir_set"], tb["array_set"]
    ref_arr = rh.make_miro_filter_set(np.zeros((1, C, B), np.float32), ar._irs_grid.azimuth,
                                      ar._irs_grid.colatitude, ar._irs_grid.weight, order, False, B)
    ref_hrir = rh.make_miro_filter_set(hr._irs_td[:, :, :128].copy(), hr._irs_grid.azimuth,
                                       hr._irs_grid.colatitude, None, order, True, B)
    ref_hrir.calculate_filter_blocks_fd(B)
    r_ref._convolver = rh.make_sh_convolver(ref_hrir, ref_arr, B, tracker)
    y_ref = session(r_ref)

    _swap_in_repo_classes(monkeypatch, jr)
    r_new = _alive(jr.JackRenderer(**kw))
    conv = Convolver.create_instance_by_filter_set(hr, B, [(0, 0)], tracker)
    conv.prepare_sh_processing(tb["sh_config"], 0, None, comp_nm=tb["comp_nm"])
    r_new._convolver = conv
    assert type(r_new._convolver) is jr.AdjustableShConvolver  # the class identity callers test
    y_new = session(r_new)
    err = np.abs(y_new - y_ref).reshape(T, -1).max(axis=1)
    assert err.max() <= TOL, err
    print("reference JackRenderer with the CUDA SH renderer: max abs err", err.max(), "peak", np.abs(y_ref).max())
