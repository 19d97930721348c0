"""Condenses an `ncu --page raw --csv` dump into one line per captured launch (the numbers quoted in DESIGN.md)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]


def col(name):
    for i, h in enumerate(hdr):
        if h == name:
            return i
    return None


KEYS = [('dur_us', 'gpu__time_duration.sum'), ('dram_rd_MB', 'dram__bytes_read.sum'), ('dram_wr_MB', 'dram__bytes_write.sum'),
        ('dram_pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
        ('tensor_pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
        ('tensor_pct_rt', 'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'),
        ('bf16_ops_pct', 'sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed'),
        ('warps_active_pct', 'sm__warps_active.avg.pct_of_peak_sustained_active'),
        ('regs', 'launch__registers_per_thread'), ('grid', 'launch__grid_size'), ('smem_dyn_KB', 'launch__shared_mem_per_block_dynamic'),
        ('l2_hit_pct', 'lts__t_sector_hit_rate.pct')]


def conv(v, u):
    try:
        x = float(v.replace(',', ''))
    except ValueError:
        return v
    scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6, 'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u)
    return round(x * scale, 3) if scale else round(x, 2)


print('\t'.join(['kernel'] + [k for k, _ in KEYS]))
for r in rows[2:]:
    name = r[col('Kernel Name')].split('(')[0].replace('void ', '').replace('pu::', '')[:44]
    out = [name]
    for k, m in KEYS:
        c = col(m)
        out.append(str(conv(r[c], units[c])) if c is not None else 'NA')
    print('\t'.join(out))
