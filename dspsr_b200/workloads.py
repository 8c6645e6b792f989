"""The configurations of BASELINE.json, as plain dictionaries."""
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
VELA_POLYCO = os.path.join(os.path.dirname(_HERE), "tests", "golden", "vela.polyco")

CFG1 = dict(
    name="cfg1: Benchmark/bench.csh CASPSR 8-bit dual-pol 400 MHz real, dspsr -F 256:D, DM 67.99, "
         "Coherence, fold 1024 bins (vela.polyco)",
    format="CASPSR8", input_real=True, input_nchan=1, npol=2, nbit=8,
    freq=1382.0, bw=-400.0, tsamp_us=0.00125, utc_start="2010-04-13-02:05:45",
    nchan=256, dm=67.99, nbin=1024, state="Coherence", ndim=4, filterbank=True,
    expect=dict(freq_res=8192, nfilt_pos=457, nfilt_neg=459, nsamp_fft=4194304, nsamp_step=3725312),
)

CFG2 = dict(
    name="cfg2: digifil-style 2-bit dual-pol 128 MHz, -F 4096:D, DM 50, total intensity, no fold "
         "(8-bit stand-in input until the 2-bit excision unpacker lands)",
    format="GENERIC8", input_real=True, input_nchan=1, npol=2, nbit=8,
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


_INSTRUMENT = {"CASPSR8": "CASPSR", "GENERIC8": "UNKNOWN", "MEERKAT8": "MKBF", "UWB16": "UWB", "CPSR2": "CPSR2"}


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
