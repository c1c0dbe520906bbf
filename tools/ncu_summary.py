"""Condense an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into the per-launch metrics DESIGN.md / bench.py cite."""
import csv
import subprocess
import sys

KEEP = [
    ('Kernel Name', 'kernel'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'regs'),
    ('gpu__time_duration.sum', 'duration_us'), ('dram__bytes_read.sum', 'dram_read'), ('dram__bytes_write.sum', 'dram_write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'), ('lts__t_bytes.sum', 'l2_bytes'),
    ('l1tex__m_xbar2l1tex_read_bytes.sum', 'l2_to_sm_read'), ('l1tex__m_xbar2l1tex_read_bytes.sum.per_second', 'l2_to_sm_rate'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_pct'),
    ('sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active', 'tensor_hmma_pct_active'),
    ('sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_hmma_cycles_pct'),
    ('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_subpipe_hmma_pct'),
    ('launch__occupancy_limit_registers', 'occ_limit_regs'), ('launch__cluster_dim_x', 'cluster_x'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'),
    ('TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor_pipe_pct'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
    ('smsp__inst_executed.sum', 'instructions'), ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_bank_conflicts'),
    ('launch__shared_mem_per_block_dynamic', 'dyn_smem'), ('launch__occupancy_limit_shared_mem', 'occ_limit_smem'),
]


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(src, dst) for src, dst in KEEP if src in idx]
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow([dst + (f' [{units[idx[src]]}]' if units[idx[src]] else '') for src, dst in cols])
        for r in rows[2:]:
            w.writerow([r[idx[src]][:80] for src, _ in cols])
    print('wrote', out, len(rows) - 2, 'launches')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
