"""raw ncu csv (ncu -i x.ncu-rep --page raw --csv) -> the key-counter summary committed under profiles/.

    python tools/ncu_summary.py gpurun_out/r2_full_pk4_raw.csv profiles/r2_ncu_full_v1_summary.csv
"""
import csv
import sys

COLS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed_op_tma_ld.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size"]


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = rows[0]
    idx = [hdr.index(c) for c in COLS if c in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[i].replace("void ", "").replace("bfm::", "") if k == 0 else r[i] for k, i in enumerate(idx)])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
