"""Key metrics of an `ncu --set full` capture (raw page, CSV) per kernel, units normalised (dev tool).
usage: python scripts/ncu_summary.py <raw.csv> [...]"""
import csv, sys
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}
KEYS = [("gpu__time_duration.sum", "duration_us"), ("dram__bytes_read.sum", "dram_read_B"), ("dram__bytes_write.sum", "dram_write_B"),
        ("lts__t_bytes.sum", "l2_bytes"), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
        ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "dmma_pipe_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_slots_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("smsp__inst_executed.sum", "warp_insts")]
STALLS = "smsp__pcsamp_warps_issue_stalled_"
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, r, units)}
        out = []
        for k, name in KEYS:
            if k in d:
                v, u = d[k]
                try:
                    x = float(v.replace(",", "")) * UNIT.get(u, 1.0)
                    out.append("%s=%s" % (name, ("%.0f" % x) if x >= 100 else ("%.2f" % x)))
                except ValueError:
                    out.append("%s=%s%s" % (name, v, u))
        st = {h[len(STALLS):]: float(d[h][0].replace(",", "")) for h in d if h.startswith(STALLS) and "not_issued" not in h and d[h][0] not in ("", "n/a")}
        tot = sum(st.values()) or 1.0
        top = ", ".join("%s %.0f%%" % (k, 100 * v / tot) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:5])
        print("%s: %s\n    %s\n    stalls: %s" % (path.split("/")[-1], r[ki][:70], " ".join(out), top))
