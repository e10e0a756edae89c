"""Summarise ncu outputs (launch list csv, full .ncu-rep) into text for profiles/."""
import collections, csv, subprocess, sys

def launch_list(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]; ki, mv, mn = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[mn] != 'gpu__time_duration.sum': continue
        name = r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[:70]
        agg.setdefault(name, []).append(float(r[mv].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    out = [f"{'kernel':70s} launches   total_us  share   avg_us"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"{k:70s} {len(v):5d} {sum(v)/1e3:11.1f} {100*sum(v)/tot:6.1f}% {sum(v)/len(v)/1e3:9.1f}")
    return "\n".join(out)

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__icc_request_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct"]

def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index('Kernel Name'); out = []
    for r in data:
        out.append("== " + r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', ''))
        for w in WANT:
            if w in hdr: out.append(f"   {w:75s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}")
        st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[i])) for i, h in enumerate(hdr)
              if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h]
        out.append("   stall cycles per issued instruction: " + ", ".join(f"{n} {v:.2f}" for n, v in sorted(st, key=lambda x: -x[1])[:7]))
    return "\n".join(out)

if __name__ == "__main__":
    kind, path = sys.argv[1], sys.argv[2]
    print(launch_list(path) if kind == "launches" else full(path))
