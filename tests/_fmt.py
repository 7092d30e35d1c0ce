"""Text formats of the reference CLI (src/cfunc.c) for comparing against its stdout."""
import numpy as np


def f6(x) -> str:
    return "%f" % float(x)  # printf("%f") of a float promoted to double


def event_long(read_id, start, length, mean, stdv) -> str:
    """print_events(), long form (cfunc.c:51-58): one line per event, then a blank line."""
    out = []
    for j in range(len(start)):
        s = int(start[j])
        out.append("%s\t%d\t%d\t%d\t%s\t%s\n" % (read_id, j, s, s + int(length[j]), f6(mean[j]), f6(stdv[j])))
    out.append("\n")
    return "".join(out)


def event_compact(read_id, n_samples, start, length) -> str:
    """print_events(), compact form (cfunc.c:19-49)."""
    n = len(start)
    s = "%s\t%d\t" % (read_id, n_samples)
    s += "%d\t%d\t" % (int(start[0]), int(start[-1]) + int(length[-1]))
    s += "%d\t" % n
    parts = []
    for j in range(n):
        mi = int(length[j])
        if mi:
            parts.append(("%d," if j < n - 1 else "%d") % mi)
    return s + "".join(parts) + "\n"


def pa_line(read_id, pa) -> str:
    """pa_func (cfunc.c:85-102)"""
    return "%s\t%d\t" % (read_id, len(pa)) + ",".join(f6(v) for v in pa) + "\n"


def stat_line(read_id, n, st) -> str:
    """stat_func (cfunc.c:126-159): note the trailing tab"""
    return "%s\t%d\t%s\t%s\t%s\t%s\t%d\t%s\t\n" % (read_id, n, f6(st[0]), f6(st[1]), f6(st[2]), f6(st[3]),
                                                    int(st[4]), f6(st[5]))


def ent_line(read_id, e3) -> str:
    """entmain's per-record line (ent.c:108-163): three "%f" doubles"""
    return "%s\t%s\t%s\t%s\n" % (read_id, f6(e3[0]), f6(e3[1]), f6(e3[2]))


def jnn_line(read_id, n, segs, compact=False, have_signal=True) -> str:
    """jnn_func (cfunc.c:108-117) + jnn_print (jnn.c:303-343)"""
    s = "%s\t%d\t" % (read_id, n)
    if have_signal:
        s += "%d\t" % len(segs)
        if compact:
            ci = 0
            for x, y in segs:
                mi = (int(x) - ci) & 0xFFFFFFFFFFFFFFFF
                ci = (ci + mi) & 0xFFFFFFFFFFFFFFFF
                if mi:
                    s += "%dH" % _i32(mi)
                mi = (int(y) - ci) & 0xFFFFFFFFFFFFFFFF
                ci = (ci + mi) & 0xFFFFFFFFFFFFFFFF
                if mi:
                    s += "%d," % _i32(mi)
        else:
            s += "".join("%d,%d;" % (int(x), int(y)) for x, y in segs)
        if len(segs) == 0:
            s += "."
    return s + "\n"


def _i32(u):
    u &= 0xFFFFFFFF
    return u - (1 << 32) if u >= (1 << 31) else u


def prefix_line(read_id, n, pos, st, p_stat=False) -> str:
    """prefix_func (cfunc.c:169-234)"""
    s = "%s\t%d\t" % (read_id, n)
    if pos[1] > 0:
        s += "%d\t%d\t" % (pos[0], pos[1])
        s += "%d\t%d" % (pos[2] + pos[1], pos[3] + pos[1]) if pos[3] > 0 else ".\t."
        if p_stat:
            s += "\t%s\t%s\t%s\t" % (f6(st[0]), f6(st[1]), f6(st[2]))
            s += "\t%s\t%s\t%s\t" % (f6(st[3]), f6(st[4]), f6(st[5])) if pos[3] > 0 else "\t.\t.\t."
    else:
        s += ".\t.\t.\t."
    return s + "\n"


PREFIX_HDR = "read_id\tlen_raw_signal\tadapt_start\tadapt_end\tpolya_start\tpolya_end"
PREFIX_HDR_STAT = "\tadapt_mean\tadapt_std\tadapt_median\tpolya_mean\tpolya_std\tpolya_median"
JNN_HDR = "read_id\tlen_raw_signal\tnum_seg\tseg\n"
ENT_HDR = "read_id\traw_ent\tdelta_ent\tbyte_ent\n"
EVENT_HDR_LONG = "read_id\tevent_idx\traw_start\traw_end\tevent_mean\tevent_std\n"
EVENT_HDR_COMPACT = "read_id\tlen_raw_signal\traw_start\traw_end\tnum_event\tevents\n"
PA_HDR = "read_id\tlen_raw_signal\tpa\n"
STAT_HDR = "read_id\tlen_raw_signal\traw_mean\tpa_mean\traw_std\tpa_std\traw_median\tpa_median\n"


def load_npz(path):
    """-> list of (read_id, (raw, digitisation, offset, range))"""
    d = np.load(path)
    off = d["read_off"]
    out = []
    for r in range(len(d["read_ids"])):
        raw = d["samples"][int(off[r]):int(off[r + 1])]
        out.append((str(d["read_ids"][r]), (raw, float(d["digitisation"][r]), float(d["offset"][r]),
                                            float(d["range"][r]))))
    return out
