"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's per-block rendering chain.

This is the parity ORACLE for the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; nothing under
``retisar_b200/`` does.  Every function cites the reference lines (relative to
``/root/reference/ReTiSAR``) it restates.  The restatement deliberately performs the SAME numpy
calls in the same order as the reference (``np.fft.rfft`` on float32 -> complex64 on numpy 2,
``np.dot``, ``np.sum(axis=0)``, float16 yaw arithmetic), so that timing it is representative of
the reference's CPU cost and its results are identical to the reference's to the last bit.

Parity pin: ``tests/test_oracle_vs_golden.py`` checks these classes against
``tests/golden/*.npz``, which ``oracle/make_golden.py`` produced by running the UNMODIFIED
reference classes (via ``oracle/ref_harness.py``) in the build container.

Third-party arithmetic restated (absent from /root/reference): ``sound_field_analysis``
(>=2021.2.4, ``setup.py:51``) functions ``process.spatFT_RT``, ``sph.mnArrays``,
``sph.reverseMnIds``, ``sph.sph_harm_all``.
Row a-X (partitioned SH-domain HRIR) does not exist in the reference: PARITY UNPINNED for that
row; it is defined by composition of the reference's own partition-queue semantics.
"""
import numpy as np

AZIM = 3  # HeadTracker.DataIndex.AZIM, `_tracker.py:73-81`


# --------------------------------------------------------------------------------------------
# spherical harmonics helpers (sound_field_analysis, restated)
# --------------------------------------------------------------------------------------------
def mn_arrays(n_max):
    """`sfa.sph.mnArrays` (call site `_filter_set.py:1259`): m = [0,-1,0,1,-2,..], n = [0,1,1,1,..]."""
    m = np.concatenate([np.arange(-n, n + 1) for n in range(n_max + 1)])
    n = np.concatenate([np.full(2 * n + 1, n) for n in range(n_max + 1)])
    return m, n


def reverse_mn_ids(n_max):
    """`sfa.sph.reverseMnIds` (call site `_convolver.py:1095`): index of (n,-m) for every (n,m)."""
    ids = []
    for n in range(n_max + 1):
        ids.extend(range(n * n + 2 * n, n * n - 1, -1))
    return ids


def sph_harm_all(n_max, az, co):
    """`sfa.sph.sph_harm_all` (call site `_filter_set.py:1263`): Y_n^m on [points; (N+1)^2]."""
    from scipy.special import sph_harm_y

    m, n = mn_arrays(n_max)
    mA, azA = np.meshgrid(m, np.asarray(az, dtype=np.float64).ravel())
    nA, coA = np.meshgrid(n, np.asarray(co, dtype=np.float64).ravel())
    return sph_harm_y(nA, mA, coA, azA)


def sh_configuration(n_max, azimuth, colatitude, weight, dtype=np.float32, enforce_pinv=False):
    """`FilterSetMiro.get_sh_configuration` (`_filter_set.py:1237-1280`) -> (sh_m, Y_w[K,C])."""
    cdtype = np.complex64 if dtype == np.float32 else np.complex128
    azimuth = np.asarray(azimuth, dtype=dtype)
    colatitude = np.asarray(colatitude, dtype=dtype)
    sh_m = mn_arrays(n_max)[0].astype(np.int16)
    sh_bases = sph_harm_all(n_max, azimuth, colatitude).astype(cdtype)
    if enforce_pinv or weight is None:
        sh_bases_weighted = np.linalg.pinv(sh_bases)
    else:
        sh_bases_weighted = np.conj(sh_bases).T * (4 * np.pi * np.asarray(weight, dtype=dtype))
    return sh_m, sh_bases_weighted


# --------------------------------------------------------------------------------------------
# filter precompute (L0)
# --------------------------------------------------------------------------------------------
def zero_pad(irs_td, block_length):
    """`FilterSet._zero_pad` (`_filter_set.py:599-623`)."""
    if not block_length:
        return irs_td
    block_count = irs_td.shape[2] / block_length
    if block_count == int(block_count) and irs_td.shape[-1] >= block_length:
        return irs_td
    block_count = int(np.ceil(block_count))
    padded = np.zeros((irs_td.shape[0], irs_td.shape[1], block_length * block_count))
    padded[:, :, : irs_td.shape[2]] = irs_td
    return padded.astype(irs_td.dtype)


def filter_blocks_fd(irs_td, block_length):
    """`FilterSet.calculate_filter_blocks_fd` (`_filter_set.py:625-664`) -> [P, dirs, E, F]."""
    if not block_length:
        block_length = irs_td.shape[-1]
    block_count = irs_td.shape[-1] // block_length
    blocks = np.stack(
        [np.fft.rfft(b, block_length * 2) for b in np.dsplit(irs_td, block_count)]
    )
    if irs_td.dtype == np.float32:
        blocks = blocks.astype(np.complex64)
    return blocks


def filter_blocks_nm(irs_blocks_fd, sh_bases_weighted):
    """`FilterSetMiro.calculate_filter_blocks_nm` (`_filter_set.py:1093-1149`) -> [P, K, E, F].

    The reference raises NotImplementedError for P > 1 (`:1136-1140`); the restatement applies
    the same per-block transform to every partition (needed by row a-X).
    """
    out = []
    for block_fd in irs_blocks_fd:
        ir_nm = np.stack(
            [np.dot(sh_bases_weighted, block_fd[:, ch, :]) for ch in range(block_fd.shape[1])]
        )
        out.append(np.swapaxes(ir_nm.astype(block_fd.dtype), 0, 1))
    return np.stack(out).copy()


def crossfade_windows(block_length, dtype=np.float32):
    """cos^2 windows of `AdjustableFdConvolver.__init__` (`_convolver.py:741-747`)."""
    w_out = np.arange(block_length, dtype=dtype)
    w_out = np.square(np.cos(w_out * np.pi / (2 * (block_length - 1))))
    w_in = np.flip(w_out).copy()
    return w_in, w_out


# --------------------------------------------------------------------------------------------
# a4: yaw (degrees, float32 in the tracker array) -> float16-quantised azimuth
# --------------------------------------------------------------------------------------------
def individual_azimuth_deg(tracker_azim, sources_deg16, is_hrir=True):
    """`AdjustableFdConvolver._calculate_individual_directions` (`_convolver.py:947-975`).

    `sources_deg16` is the float16 array of `_convolver.py:729`; the python-float tracker value
    is a weak scalar, so all arithmetic happens in float16 exactly as in the reference.
    """
    tracker_dir = 1 if is_hrir else -1
    azims_deg = tracker_dir * float(tracker_azim) + sources_deg16[:, 0]
    return (azims_deg + 360.0) % 360.0


def sh_rotation_phases(azim_deg16, sh_m, cdtype=np.complex64):
    """`_convolver.py:1279-1281`: exp(-1j * m * deg2rad(az)) with az a float16 array of size 1."""
    return np.exp(cdtype(-1j) * sh_m * np.deg2rad(azim_deg16))


# --------------------------------------------------------------------------------------------
# a1, a6, a8: uniformly partitioned overlap-save convolver
# --------------------------------------------------------------------------------------------
class OlsOracle:
    """`OverlapSaveConvolver` (`_convolver.py:436-669`), channel-wise FIR (ARIR / HPCF / benchmark).

    irs_blocks_fd: [P, 1, Cout, F] complex (as `FilterSetMultiChannel.get_filter_blocks_fd`,
    `_filter_set.py:799-823`).
    """

    def __init__(self, irs_blocks_fd, block_length):
        self.B = block_length
        self.H = irs_blocks_fd
        self.blocks_fd = np.zeros_like(irs_blocks_fd[:, :1])  # `:470-472` via dirac blocks
        rdtype = np.float32 if irs_blocks_fd.dtype == np.complex64 else np.float64
        self.input_block_td = np.zeros((self.blocks_fd.shape[-2], 2 * self.B), dtype=rdtype)
        self.is_passthrough = False

    def _shift_and_convert_input(self, x):  # `:573-600`
        n = x.shape[1]
        self.input_block_td[:, :n] = self.input_block_td[:, n:]
        self.input_block_td[:, n:] = x  # broadcasts a mono input into all rows
        return np.fft.rfft(self.input_block_td)

    def _shift_and_convert_result(self, blocks_fd):  # `:624-669`
        first_td = np.fft.irfft(blocks_fd[0])
        if blocks_fd.shape[0] > 1:
            blocks_fd = np.roll(blocks_fd, -1, axis=0)
        blocks_fd[-1] = 0.0
        return blocks_fd, first_td[0, :, first_td.shape[-1] // 2:]

    def clear(self):  # `:500-505`
        self.input_block_td.fill(0)
        self.blocks_fd.fill(0)

    def filter_block(self, x):  # `:524-571`
        x_fd = self._shift_and_convert_input(x)
        if self.is_passthrough:
            x_fd = x_fd[: self.blocks_fd.shape[-2]]
            self.blocks_fd[0, 0, : x_fd.shape[-2]] = x_fd
        else:
            for block_fd, filter_fd in zip(self.blocks_fd, self.H):  # `:602-622`
                block_fd += filter_fd * x_fd
        self.blocks_fd, y = self._shift_and_convert_result(self.blocks_fd)
        return y


# --------------------------------------------------------------------------------------------
# a1-a7, a10: SH-domain renderer
# --------------------------------------------------------------------------------------------
class ShOracle:
    """`AdjustableShConvolver` (`_convolver.py:978-1341`), one render stream.

    sh_m [K] int16, sh_bases_weighted [K, C] complex, irs_blocks_nm [1, K, E, F] complex
    (HRIR SH coefficients x compensation spectra, `_convolver.py:1189-1206`).
    """

    def __init__(self, sh_m, sh_bases_weighted, irs_blocks_nm, block_length, sh_max_order,
                 source_azim_deg=0.0, is_hrir=True):
        self.B = block_length
        cdtype = irs_blocks_nm.dtype
        rdtype = np.float32 if cdtype == np.complex64 else np.float64
        self.sh_m = np.asarray(sh_m).copy()
        self.rev = reverse_mn_ids(sh_max_order)
        self.sh_max_order = sh_max_order
        self.sh_cur_order = sh_max_order
        self.Yw = sh_bases_weighted.copy()
        self.H = irs_blocks_nm
        E, F = irs_blocks_nm.shape[-2:]
        self.blocks_fd = np.zeros((1, 1, E, F), dtype=cdtype)  # `:470,:732`
        self.last_blocks_fd = np.zeros_like(self.blocks_fd)  # `:750`
        self.last_sh_azim_nm = np.zeros((self.Yw.shape[0], 1, 1), dtype=self.Yw.dtype)  # `:1098`
        self.input_block_td = np.zeros((self.Yw.shape[-1], 2 * self.B), dtype=rdtype)  # `:1113`
        self.sources_deg = np.array([(source_azim_deg, 0)], dtype=np.float16)  # `:729`
        self.is_hrir = is_hrir
        self.is_crossfade = True  # `:722`
        self.is_passthrough = False
        self.w_in, self.w_out = crossfade_windows(self.B, rdtype)

    def set_crossfade(self, state):  # `:868-892`, `:1318-1341`
        self.is_crossfade = state
        if not state:
            self.last_blocks_fd.fill(0)
            self.last_sh_azim_nm.fill(0)

    def set_order(self, order):  # `:1159-1160` (compensation re-application is setup-time)
        self.sh_cur_order = order

    def clear(self):  # `:774-778`
        self.input_block_td.fill(0)
        self.blocks_fd.fill(0)
        self.last_blocks_fd.fill(0)

    def _irfft_second_half(self, blocks_fd):  # `:645-668` with P == 1
        first_td = np.fft.irfft(blocks_fd[0])
        blocks_fd[-1] = 0.0
        return first_td[0, :, first_td.shape[-1] // 2:]

    def filter_block(self, x, yaw_deg):
        B = self.B
        # a1 `:573-600`
        self.input_block_td[:, :B] = self.input_block_td[:, B:]
        self.input_block_td[:, B:] = x
        x_fd = np.fft.rfft(self.input_block_td)

        if self.is_passthrough:  # a10 `:1256-1257` -> `:558-562`
            x_fd = x_fd[: self.blocks_fd.shape[-2]]
            self.blocks_fd[0, 0, : x_fd.shape[-2]] = x_fd
            return self._irfft_second_half(self.blocks_fd)

        # a2 `:1260-1263`
        nm = np.dot(self.Yw, x_fd)
        # a3 `:1272-1275`
        nm = nm[self.rev][:, np.newaxis, :]
        nm = np.repeat(nm, self.blocks_fd.shape[-2], axis=1)
        nm *= self.H[0]
        # a4 `:1278-1283`
        azim = individual_azimuth_deg(yaw_deg, self.sources_deg, self.is_hrir)
        sh_azim = sh_rotation_phases(azim, self.sh_m, self.blocks_fd.dtype.type)
        sh_azim = sh_azim[: (self.sh_cur_order + 1) ** 2, np.newaxis, np.newaxis]
        # a5 + a6 `:1287-1293`
        self.blocks_fd[0, 0] = np.sum(nm[: sh_azim.shape[0]] * sh_azim, axis=0)
        y_in = self._irfft_second_half(self.blocks_fd)
        if not self.is_crossfade:  # `:1296-1297`
            return y_in
        # `:1301-1308`
        self.last_blocks_fd[0, 0] = np.sum(
            nm[: self.last_sh_azim_nm.shape[0]] * self.last_sh_azim_nm, axis=0
        )
        y_out = self._irfft_second_half(self.last_blocks_fd)
        self.last_sh_azim_nm = sh_azim  # `:1311`
        return (y_in * self.w_in) + (y_out * self.w_out)  # a7 `:1314-1316`


# --------------------------------------------------------------------------------------------
# a-X: partitioned SH-domain HRIR (EXTENSION, parity unpinned by the reference)
# --------------------------------------------------------------------------------------------
class ShPartitionedOracle(ShOracle):
    """SH renderer with a P-partition HRIR, by composition with the queue semantics of
    `AdjustableFdConvolver.filter_block` (`_convolver.py:802-836`) and
    `_filter_block_shift_and_convert_result` (`:624-669`): the current rotation is applied at
    input time into an output-side queue, a second queue receives the previous rotation, both
    queues are rolled every block and their heads cross-faded.  For P == 1 this is ShOracle.
    """

    def __init__(self, sh_m, sh_bases_weighted, irs_blocks_nm, block_length, sh_max_order, **kw):
        super().__init__(sh_m, sh_bases_weighted, irs_blocks_nm, block_length, sh_max_order, **kw)
        P = irs_blocks_nm.shape[0]
        E, F = irs_blocks_nm.shape[-2:]
        self.blocks_fd = np.zeros((P, 1, E, F), dtype=irs_blocks_nm.dtype)
        self.last_blocks_fd = np.zeros_like(self.blocks_fd)

    def _roll(self, blocks_fd):
        first_td = np.fft.irfft(blocks_fd[0])
        if blocks_fd.shape[0] > 1:
            blocks_fd = np.roll(blocks_fd, -1, axis=0)
        blocks_fd[-1] = 0.0
        return blocks_fd, first_td[0, :, first_td.shape[-1] // 2:]

    def filter_block(self, x, yaw_deg):
        B = self.B
        self.input_block_td[:, :B] = self.input_block_td[:, B:]
        self.input_block_td[:, B:] = x
        x_fd = np.fft.rfft(self.input_block_td)
        if self.is_passthrough:  # `_convolver.py:558-562` then the queue roll `:645-665`
            x_fd = x_fd[: self.blocks_fd.shape[-2]]
            self.blocks_fd[0, 0, : x_fd.shape[-2]] = x_fd
            self.blocks_fd, y = self._roll(self.blocks_fd)
            return y
        nm = np.dot(self.Yw, x_fd)[self.rev][:, np.newaxis, :]
        azim = individual_azimuth_deg(yaw_deg, self.sources_deg, self.is_hrir)
        sh_azim = sh_rotation_phases(azim, self.sh_m, self.blocks_fd.dtype.type)
        sh_azim = sh_azim[: (self.sh_cur_order + 1) ** 2, np.newaxis, np.newaxis]
        n_cur = sh_azim.shape[0]
        for p in range(self.H.shape[0]):
            self.blocks_fd[p, 0] += np.sum(nm[:n_cur] * self.H[p, :n_cur] * sh_azim, axis=0)
        self.blocks_fd, y_in = self._roll(self.blocks_fd)
        if not self.is_crossfade:
            return y_in
        n_last = self.last_sh_azim_nm.shape[0]
        for p in range(self.H.shape[0]):
            self.last_blocks_fd[p, 0] += np.sum(
                nm[:n_last] * self.H[p, :n_last] * self.last_sh_azim_nm, axis=0
            )
        self.last_blocks_fd, y_out = self._roll(self.last_blocks_fd)
        self.last_sh_azim_nm = sh_azim
        return (y_in * self.w_in) + (y_out * self.w_out)


# --------------------------------------------------------------------------------------------
# f3: HRIR / BRIR switching renderer
# --------------------------------------------------------------------------------------------
class FdOracle:
    """`AdjustableFdConvolver` (`_convolver.py:672-975`) with a `FilterSetSsr`
    (`_filter_set.py:843-899`): per source the filter of the direction `int(-azimuth)` is applied
    at input time into a P-partition output queue, a second queue receives the previous block's
    filters, heads are cross-faded.

    irs_blocks_fd: [P, directions, 2, F] complex.
    """

    def __init__(self, irs_blocks_fd, block_length, sources_deg, is_hrir=True):
        self.B = block_length
        self.H = irs_blocks_fd
        P, _, E, F = irs_blocks_fd.shape
        cdtype = irs_blocks_fd.dtype
        rdtype = np.float32 if cdtype == np.complex64 else np.float64
        self.sources_deg = np.array(sources_deg, dtype=np.float16)  # `:729`
        n_src = self.sources_deg.shape[0]
        self.blocks_fd = np.zeros((P, 1, E, F), dtype=cdtype)  # `:732`
        self.last_blocks_fd = np.zeros_like(self.blocks_fd)  # `:750`
        self.last_filters_fd = np.zeros((1, P, E, F), dtype=cdtype)  # `:751` (dirac-shaped)
        self.input_block_td = np.zeros((n_src, 2 * self.B), dtype=rdtype)  # `:735-738`
        self.is_hrir = is_hrir
        self.is_crossfade = True
        self.is_passthrough = False
        self.w_in, self.w_out = crossfade_windows(self.B, rdtype)

    def set_crossfade(self, state):  # `:868-892`
        self.is_crossfade = state
        if not state:
            self.last_blocks_fd.fill(0)
            self.last_filters_fd.fill(0)

    def clear(self):  # `:774-778`
        self.input_block_td.fill(0)
        self.blocks_fd.fill(0)
        self.last_blocks_fd.fill(0)

    def _current_filters(self, yaw_deg):  # `:913-945`
        tracker_dir = 1 if self.is_hrir else -1
        azims = (tracker_dir * float(yaw_deg) + self.sources_deg[:, 0] + 360.0) % 360.0
        return np.stack([self.H[:, int(-a)] for a in azims])  # `_filter_set.py:898-899`

    @staticmethod
    def _multiply(blocks_fd, filters_fd, x_fd):  # `:838-866`
        for s in range(filters_fd.shape[0]):
            for block_fd, filter_fd in zip(blocks_fd, filters_fd[s]):
                block_fd[0] += filter_fd * x_fd[s] / filters_fd.shape[0]

    @staticmethod
    def _roll(blocks_fd):  # `:624-669`
        first_td = np.fft.irfft(blocks_fd[0])
        if blocks_fd.shape[0] > 1:
            blocks_fd = np.roll(blocks_fd, -1, axis=0)
        blocks_fd[-1] = 0.0
        return blocks_fd, first_td[0, :, first_td.shape[-1] // 2:]

    def filter_block(self, x, yaw_deg):  # `:780-836`
        B = self.B
        self.input_block_td[:, :B] = self.input_block_td[:, B:]
        self.input_block_td[:, B:] = x
        x_fd = np.fft.rfft(self.input_block_td)
        if self.is_passthrough:  # `:799-800` -> `:558-562`
            x_fd = x_fd[: self.blocks_fd.shape[-2]]
            self.blocks_fd[0, 0, : x_fd.shape[-2]] = x_fd
            self.blocks_fd, y = self._roll(self.blocks_fd)
            return y
        filters_fd = self._current_filters(yaw_deg)
        self._multiply(self.blocks_fd, filters_fd, x_fd)
        self.blocks_fd, y_in = self._roll(self.blocks_fd)
        if not self.is_crossfade:
            return y_in
        self._multiply(self.last_blocks_fd, self.last_filters_fd, x_fd)
        self.last_blocks_fd, y_out = self._roll(self.last_blocks_fd)
        self.last_filters_fd = filters_fd
        return (y_in * self.w_in) + (y_out * self.w_out)


# --------------------------------------------------------------------------------------------
# f4: block delay line of the JACK clients
# --------------------------------------------------------------------------------------------
class DelayBufferOracle:
    """`DelayBuffer` (`_convolver.py:11-141`): ring of blocks, delay in whole blocks."""

    def __init__(self, sample_rate, block_length, delay_ms=0.0):  # `:33-50`
        self.buffer = None
        self.ptr = 0
        self.fs, self.B = sample_rate, block_length
        self.delay_blocks = self._blocks(delay_ms)

    def init(self, max_length_sec, channel_count):  # `:52-67`
        n = self._blocks(max_length_sec * 1000) + 1
        self.buffer = np.zeros((n, channel_count, self.B), dtype=np.float32)
        self.ptr = 0

    def _blocks(self, delay_ms):  # `:69-81`
        return int(np.ceil((delay_ms / 1000) * self.fs / self.B))

    def set_delay(self, new_delay_ms=0.0):  # `:89-115`
        if self.buffer is None:
            return self.delay_blocks
        self.delay_blocks = min(self._blocks(new_delay_ms), self.buffer.shape[0])
        if self.delay_blocks == 0:
            self.buffer.fill(0)
            self.ptr = 0
        return (self.delay_blocks * self.B / self.fs) * 1000

    def process_block(self, x):  # `:117-141`
        if self.buffer is None or not self.delay_blocks or x is None:
            return x
        self.buffer[self.ptr] = x
        y = self.buffer[self.ptr - self.delay_blocks]  # python index: -ring size wraps to 0
        self.ptr = (self.ptr + 1) % self.buffer.shape[0]
        return y


def calculate_rms(data, is_level=False):
    """`tools.calculate_rms` (`tools.py:1117-1145`), real data."""
    rms = np.sqrt(np.mean(np.square(np.abs(data)), axis=-1))
    if is_level:
        rms[np.nonzero(rms == 0)] = np.nan
        with np.errstate(divide="ignore"):
            rms = 20 * np.log10(rms)
        rms[np.isnan(rms)] = -200
    return rms


def calculate_peak(data_td, is_level=False):
    """`tools.calculate_peak` (`tools.py:1148-1169`)."""
    peak = np.nanmax(np.abs(data_td), axis=-1)
    if is_level:
        peak[np.nonzero(peak == 0)] = np.nan
        with np.errstate(divide="ignore"):
            peak = 20 * np.log10(peak)
        peak[np.isnan(peak)] = -200
    return peak


def deliver(output_td, volume, volume_relative):
    """Output stage of `JackClient._process_deliver` (`_jack_client.py:789-824`): in-place volume,
    then RMS / peak levels in dBFS of the scaled block."""
    output_td *= volume * volume_relative
    return calculate_rms(output_td, is_level=True), calculate_peak(output_td, is_level=True)
