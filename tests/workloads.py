"""The configurations of BASELINE.json as plain dictionaries, and the set-up arithmetic
that bench.py's two arms and the tests share.  Pure Python + numpy: the module imports neither the product
(dspsr_b200) nor the oracle; callers pass in whichever `dedispersion` / predictor they are entitled to use."""
import math

import numpy as np

import os

_HERE = os.path.dirname(os.path.abspath(__file__))
VELA_POLYCO = os.path.join(_HERE, "golden", "vela.polyco")

CFG1 = dict(
    name="cfg1: Benchmark/bench.csh CASPSR 8-bit dual-pol 400 MHz real, dspsr -F 256:D, DM 67.99, "
         "Coherence, fold 1024 bins (vela.polyco)",
    format="CASPSR8", input_real=True, input_nchan=1, npol=2, nbit=8,
    freq=1382.0, bw=-400.0, tsamp_us=0.00125, utc_start="2010-04-13-02:05:45",
    nchan=256, dm=67.99, nbin=1024, state="Coherence", ndim=4, filterbank=True,
    expect=dict(freq_res=8192, nfilt_pos=457, nfilt_neg=459, nsamp_fft=4194304, nsamp_step=3725312),
)

CFG2 = dict(
    name="cfg2: digifil-style 2-bit dual-pol 128 MHz real (CPSR2 convention), -F 4096:D, DM 50, total intensity, no fold",
    format="TWOBIT", input_real=True, input_nchan=1, npol=2, nbit=2,
    freq=1400.0, bw=128.0, tsamp_us=0.00390625, utc_start="2010-04-13-02:05:45",
    nchan=4096, dm=50.0, nbin=0, state="Intensity", ndim=1, filterbank=True,
    expect=dict(freq_res=8, nfilt_pos=1, nfilt_neg=1, nsamp_fft=65536, nsamp_step=49152),
)

CFG3 = dict(
    name="cfg3: MeerKAT L-band 856 MHz, 1024 input channels, 8-bit dual-pol complex, DM 500, "
         "Coherence, fold 1024 bins; channel-sharded",
    format="MEERKAT8", input_real=False, input_nchan=1024, npol=2, nbit=8,
    freq=1284.0, bw=856.0, tsamp_us=1024.0 / 856.0, utc_start="2010-04-13-02:05:45",
    nchan=1024, dm=500.0, nbin=1024, state="Coherence", ndim=4, filterbank=False,
    expect=dict(freq_res=65536, nfilt_pos=2536, nfilt_neg=2543),
)

CFG4 = dict(
    name="cfg4: single-channel 400 MHz band at 12.5 GHz, DM 1500, 2^22-point overlap-save, 8-bit complex",
    format="GENERIC8", input_real=False, input_nchan=1, npol=2, nbit=8,
    freq=12500.0, bw=400.0, tsamp_us=0.0025, utc_start="2010-04-13-02:05:45",
    nchan=1, dm=1500.0, nbin=1024, state="Coherence", ndim=4, filterbank=False, nfft=4194304,
    expect=dict(freq_res=4194304, nfilt_pos=534848, nfilt_neg=588748),
)


def cfg5_subband(k):
    return dict(
        name="cfg5 sb%d: UWL-like 128 MHz sub-band at %d MHz, 16-bit dual-pol complex, -F 128:D" % (k, 768 + 128 * k),
        format="UWB16", input_real=False, input_nchan=1, npol=2, nbit=16,
        freq=768.0 + 128.0 * k, bw=128.0, tsamp_us=0.0078125, utc_start="2010-04-13-02:05:45",
        nchan=128, dm=67.99, nbin=1024, state="Coherence", ndim=4, filterbank=True,
    )


def polyco_text():
    with open(VELA_POLYCO) as f:
        return f.read()


_INSTRUMENT = {"CASPSR8": "CASPSR", "GENERIC8": "UNKNOWN", "MEERKAT8": "MKBF", "UWB16": "UWB", "CPSR2": "CPSR2", "TWOBIT": "CPSR2"}


def dada_header(cfg, obs_offset=0, hdr_size=4096):
    """4096-byte ASCII DADA header of a configuration (keys per Kernel/Classes/ASCIIObservation.C:95-400,
    SURVEY Appendix A.8), NUL-padded, so that a real dspsr could read the synthetic file."""
    lines = [
        ("HDR_VERSION", "1.0"), ("HDR_SIZE", str(hdr_size)), ("INSTRUMENT", _INSTRUMENT[cfg["format"]]),
        ("TELESCOPE", "PKS"), ("SOURCE", "J0835-4510"), ("MODE", "PSR"), ("FREQ", repr(float(cfg["freq"]))),
        ("BW", repr(float(cfg["bw"]))), ("NCHAN", str(cfg["input_nchan"])), ("NPOL", str(cfg["npol"])),
        ("NDIM", "1" if cfg["input_real"] else "2"), ("NBIT", str(cfg["nbit"])), ("TSAMP", repr(float(cfg["tsamp_us"]))),
        ("UTC_START", cfg["utc_start"]), ("OBS_OFFSET", str(obs_offset)), ("RESOLUTION", "4"),
    ]
    text = "".join("%-16s %s\n" % kv for kv in lines)
    assert len(text) < hdr_size
    return text + "\0" * (hdr_size - len(text))


def parse_dada_header(raw):
    """key -> value strings of an ASCII DADA header (first whitespace-separated token after the key, as
    ascii_header_get's sscanf does; Kernel/Classes/ascii_header.c)."""
    if isinstance(raw, bytes):
        raw = raw.split(b"\0", 1)[0].decode()
    out = {}
    for line in raw.split("\n"):
        line = line.split("#", 1)[0].split()
        if len(line) >= 2:
            out[line[0]] = line[1]
    return out


# ---------------------------------------------------------------------------------------------------------------
# set-up arithmetic shared by bench.py's arms
# ---------------------------------------------------------------------------------------------------------------
CONFIGS = {"cfg1": CFG1, "cfg2": CFG2, "cfg3": CFG3, "cfg4": CFG4}


def utc_to_mjd(utc):
    """'YYYY-MM-DD-hh:mm:ss' -> (day, second of day, fraction): the split epoch of ASCIIObservation's UTC_START."""
    y, m, d, hms = utc.split("-")
    hh, mm, ss = hms.split(":")
    y, m, d = int(y), int(m), int(d)
    a = (14 - m) // 12
    yy = y + 4800 - a
    mo = m + 12 * a - 3
    jdn = d + (153 * mo + 2) // 5 + 365 * yy + yy // 4 - yy // 100 + yy // 400 - 32045
    return (jdn - 2400001, int(hh) * 3600 + int(mm) * 60 + int(ss), 0.0)


def mjd_add(mjd, seconds):
    day, sec, frac = mjd
    frac += seconds
    whole = math.floor(frac)
    frac -= whole
    sec += int(whole)
    day += sec // 86400
    sec %= 86400
    return (day, sec, frac)


def sizes(cfg, freq_res, nfilt_pos, nfilt_neg, input_nchan=None):
    """Overlap-save bookkeeping of a configuration (Filterbank.C:107-155, Convolution.C:139-170)."""
    nin = cfg["input_nchan"] if input_nchan is None else input_nchan
    C = cfg["nchan"] // cfg["input_nchan"]             # output channels per input channel
    Nc = C * freq_res
    real = cfg["input_real"]
    nfilt = nfilt_pos + nfilt_neg
    nsamp_fft = (2 if real else 1) * Nc
    overlap = (2 if real else 1) * nfilt * C
    step = nsamp_fft - overlap
    nkeep = freq_res - nfilt
    ndim = 1 if real else 2
    rate_in = 1e6 / cfg["tsamp_us"]
    return dict(nin=nin, C=C, F=freq_res, Nc=Nc, nsamp_fft=nsamp_fft, overlap=overlap, step=step, nkeep=nkeep, ndim=ndim,
                npos=nfilt_pos, nneg=nfilt_neg, rate_in=rate_in, rate_out=rate_in * freq_res / nsamp_fft,
                bytes_per_sample=nin * cfg["npol"] * ndim * cfg["nbit"] / 8.0)


def algorithmic_bytes(cfg, S):
    """compulsory global-memory traffic per part, per polarisation, per input channel of the three-kernel
    decomposition with exactly ONE spectrum round trip: raw samples (overlap re-read included) + spectrum written +
    spectrum read [+ response read when it cannot stay in L2 (> 64 MiB)] [+ detected output when that is the product].
    Returns (bytes per part per pol per input channel, bytes per new per-pol sample)."""
    raw = S["nsamp_fft"] * cfg["nbit"] * 1 / 8.0        # nsamp_fft counts real samples (ndim folded in for complex below)
    if not cfg["input_real"]:
        raw = S["nsamp_fft"] * cfg["nbit"] * 2 / 8.0
    spec = 8.0 * S["Nc"]
    b = raw + 2 * spec
    if 8.0 * S["Nc"] * cfg["input_nchan"] > 64 * 2 ** 20:
        b += spec / cfg["npol"]
    if cfg["nbin"] == 0:
        npol_out = {"Intensity": 1, "PPQQ": 2}.get(cfg["state"], 4)
        b += 4.0 * S["C"] * S["nkeep"] * npol_out / cfg["npol"]
    return b, b / S["step"]


def algorithmic_flops(cfg, S):
    """5 Nc (log2 Nc + log2 F) + 6 Nc per part, pol and input channel (+ 10 per detected sample pair);
    returns flop per new per-pol input sample."""
    f = 5.0 * S["Nc"] * (math.log2(S["Nc"]) + math.log2(S["F"])) + 6.0 * S["Nc"]
    f += 10.0 * S["C"] * S["nkeep"] / cfg["npol"]
    return f / S["step"]


def block_phase(S, start_mjd, first_sample, phase_fn, freq_fn):
    """phi, phase_per_sample of a block whose first input sample is `first_sample`: the output starts nfilt_pos output
    samples later (Filterbank.C:370, Convolution.C:300) and Fold evaluates the predictor at the midpoint of its first
    sample (Fold.C:650-657,718-720).  phase_fn / freq_fn: (day, sec, frac) -> fractional phase / spin frequency."""
    t_block = mjd_add(start_mjd, first_sample / S["rate_in"] + S["npos"] / S["rate_out"])
    t0 = mjd_add(t_block, 0.5 / S["rate_out"])
    phi = phase_fn(t0)
    pfold = 1.0 / freq_fn(t0)
    return phi, (1.0 / S["rate_out"]) / pfold


def raw_bytes(cfg, S, ndat, seed, tile_parts=16):
    """Seeded raw stream of `ndat` NDAT-samples in the configuration's byte layout; a `tile_parts`-part random base is
    tiled to length (throughput does not depend on the data; parity is tested elsewhere on fully random streams)."""
    import synth
    res = {"CASPSR8": 4, "MEERKAT8": 256, "UWB16": 2048, "TWOBIT": 512, "GENERIC8": 1}[cfg["format"]]
    ndat = (ndat + res - 1) // res * res
    base_n = min(ndat, tile_parts * S["step"] + S["overlap"])
    base_n = (base_n + res - 1) // res * res
    fmt = cfg["format"]
    if fmt == "CASPSR8":
        base = synth.caspsr_bytes(base_n, seed=seed)
    elif fmt == "MEERKAT8":
        base = synth.meerkat_bytes(base_n, S["nin"], cfg["npol"], seed=seed)
    elif fmt == "UWB16":
        base = synth.uwb_bytes(base_n, cfg["npol"], seed=seed)
    elif fmt == "TWOBIT":
        base = synth.twobit_bytes(base_n, cfg["npol"], seed=seed)
    else:
        base = synth.generic8_bytes(base_n, S["nin"], cfg["npol"], S["ndim"], seed=seed)
    nbytes = int(round(ndat * S["bytes_per_sample"]))
    if base.size >= nbytes:
        return np.ascontiguousarray(base[:nbytes])
    # tile whole periods of the layout (the base length is a multiple of the layout resolution)
    reps = -(-nbytes // base.size)
    return np.tile(base, reps)[:nbytes].copy()


def subband_cost(S):
    """Relative cost of one input sample of a sub-band (FFT work grows with log2 Nc + log2 F)."""
    return math.log2(S["Nc"]) + math.log2(S["F"])


def assign_by_cost(costs, world):
    """Longest-processing-time assignment of items (sub-bands) to `world` ranks: -> list of item lists per rank.
    Deterministic (ties by index), every rank computes the same map ('round-robin by cost')."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(x) for x in out]


GPU_PARTS = {"cfg1": 74, "cfg2": 2048, "cfg3": 32, "cfg4": 16}      # overlap-save parts per block on one GPU
CPU_PARTS = {"cfg1": 1, "cfg2": 64, "cfg3": 1, "cfg4": 1}            # ... per block of one CPU thread
CFG3_CHAN_PER_RANK = 128                                            # 1024 channels over the 8 GPUs of a box
CFG5_BLOCK_SAMPLES = {"gpu": 1 << 24, "cpu": 1 << 19}               # samples per sub-band and block


def plan_streams(workload, rank, world, api, parts=None, target="gpu", seed=0xD5B5):
    """The pipelines ("streams") rank `rank` of `world` runs for a BASELINE configuration and how they shard
    (SURVEY 8e).  `api` supplies bittable8 / dedispersion / dedispersion_channels / predictor from whichever side the
    caller may use (the product's host maths or the oracle's).  Returns (streams, meta); a stream is a dict with the
    configuration, its overlap-save sizes S, response H, look-up table, raw bytes of one block, parts per block and
    the fold phase of that block."""
    pred = api.predictor(polyco_text())
    phase_fn, freq_fn = pred.phase, pred.frequency
    lut, lut_scale = api.bittable8()
    streams = []

    def add(cfg, S, H, npart, first_sample, chan0, sd):
        start = utc_to_mjd(cfg["utc_start"])
        phi, pps = block_phase(S, start, first_sample, phase_fn, freq_fn)
        ndat = npart * S["step"] + S["overlap"]
        streams.append(dict(cfg=cfg, S=S, H=H, lut=lut if cfg["format"] in ("CASPSR8", "GENERIC8") else None,
                            scale=np.float32(lut_scale) if cfg["format"] == "MEERKAT8" else np.float32(0),
                            parts=npart, first_sample=first_sample, chan0=chan0, phi=phi, pps=pps, ndat=ndat,
                            raw=raw_bytes(cfg, S, ndat, sd), nbin=cfg["nbin"], state=cfg["state"], dndim=cfg["ndim"]))

    if workload in ("cfg1", "cfg2", "cfg4"):
        cfg = CONFIGS[workload]
        d, H = api.dedispersion(cfg["freq"], cfg["bw"], cfg["dm"], 1, cfg["nchan"], cfg["input_real"],
                                cfg.get("nfft", 0))
        S = sizes(cfg, d.ndat, d.impulse_pos, d.impulse_neg)
        npart = parts or (GPU_PARTS if target == "gpu" else CPU_PARTS)[workload]
        add(cfg, S, H, npart, rank * npart * S["step"], 0, seed + rank)
        meta = dict(sharding="time blocks with overlap re-read (nchan=1)", combine="reduce" if cfg["nbin"] else "none",
                    rate_in=S["rate_in"], nchan_samples=1)
    elif workload == "cfg3":
        cfg = CFG3
        d, _ = api.dedispersion(cfg["freq"], cfg["bw"], cfg["dm"], cfg["input_nchan"], cfg["nchan"], False, 0, build=False)
        nloc = CFG3_CHAN_PER_RANK
        chan0 = (rank * nloc) % cfg["nchan"]
        H = api.dedispersion_channels(d, chan0, nloc)
        S = sizes(cfg, d.ndat, d.impulse_pos, d.impulse_neg, input_nchan=nloc)
        npart = parts or (GPU_PARTS if target == "gpu" else CPU_PARTS)[workload]
        add(cfg, S, H, npart, 0, chan0, seed + rank)
        meta = dict(sharding="contiguous channel ranges, %d of %d channels per GPU (weak: N GPUs = N x %d channels; "
                             "8 GPUs = the whole band)" % (nloc, cfg["nchan"], nloc),
                    combine="gather", rate_in=S["rate_in"], nchan_samples=nloc)
    elif workload == "cfg5":
        subs = []
        for k in range(26):
            cfg = cfg5_subband(k)
            d, _ = api.dedispersion(cfg["freq"], cfg["bw"], cfg["dm"], 1, cfg["nchan"], False, 0, build=False)
            subs.append((cfg, d, sizes(cfg, d.ndat, d.impulse_pos, d.impulse_neg)))
        mine = assign_by_cost([subband_cost(x[2]) for x in subs], world)[rank]
        span = parts * 1 if parts else CFG5_BLOCK_SAMPLES[target]
        for k in mine:
            cfg, d, S = subs[k]
            _, H = api.dedispersion(cfg["freq"], cfg["bw"], cfg["dm"], 1, cfg["nchan"], False, 0)
            add(cfg, S, H, max(1, span // S["step"]), 0, k, seed + k)
        meta = dict(sharding="26 sub-bands by cost over the GPUs (this rank: %s), strong scaling" % mine,
                    combine="gather", rate_in=subs[0][2]["rate_in"], nchan_samples=1, subbands=mine, strong=True)
    else:
        raise ValueError("unknown workload %r" % workload)
    return streams, meta
