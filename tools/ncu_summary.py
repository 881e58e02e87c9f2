"""Condenses an .ncu-rep (read here with `ncu -i`, no GPU needed) into the few numbers the roofline argument uses.
    python tools/ncu_summary.py gpurun_out/fmha_r1b.ncu-rep [more.ncu-rep ...] > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "sm cycles"),
    ("smsp__cycles_active.avg", "smsp cycles active"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_uniform", None),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active", "tensor (hmma subpipe) cycles active"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % of peak (active)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % of peak (elapsed)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", None),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("smsp__cycles_active.avg", None),
    ("sm__cycles_active.avg", "sm cycles active"),
    ("smsp__average_warp_latency_issue_stalled", None),
]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        print(f"== {rep}")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            print(f"kernel: {d.get('Kernel Name', '?')[:110]}")
            seen = set()
            for key, label in KEYS:
                for h in hdr:
                    if key in h and h not in seen and d.get(h, "") != "":
                        seen.add(h)
                        print(f"  {h:90s} {d[h]:>16s} {u.get(h, '')}")
            # warp stall breakdown (per issue-active sample)
            st = [(h, float(d[h])) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and d.get(h)]
            for h, v in sorted(st, key=lambda kv: -kv[1])[:8]:
                print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} {v:8.3f} warps/issue")
            print()


if __name__ == "__main__":
    main()
