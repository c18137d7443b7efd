"""Summarise an ncu report exported with --page raw / --page source --csv (development aid)."""
import csv, sys

def raw(path, want=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = want or ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
                    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
                    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
                    'launch__registers_per_thread', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
                    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'launch__grid_size',
                    'sm__cycles_elapsed.avg', 'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
                    'smsp__inst_executed_pipe_uniform.sum', 'sm__inst_executed_pipe_tensor.sum']
    for r in rows[2:]:
        print('----')
        for w in want:
            if w in idx:
                print(f"{w:72s} {r[idx[w]][:70]:>40s} {units[idx[w]]}")

def source(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr_i = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
    hdr = rows[hdr_i]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hdr_i + 1:] if len(r) >= len(hdr)]
    f = lambda r, c: float(r[idx[c]] or 0) if r[idx[c]].replace('.', '', 1).isdigit() else 0.0
    tot = sum(f(r, '# Samples') for r in data)
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    print(len(data), "instructions, samples", tot)
    agg = {c: sum(f(r, c) for r in data) for c in stalls}
    print("stall totals:", {k: int(v) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]})
    for r in sorted(sorted(data, key=lambda r: -f(r, '# Samples'))[:top], key=lambda r: r[0]):
        s = f(r, '# Samples')
        st = sorted(((f(r, c), c) for c in stalls), reverse=True)[:2]
        print(f"{r[0][-5:]} {s:7.0f} {100 * s / tot:5.1f}% exec={f(r, 'Instructions Executed'):10.0f}  {r[idx['Source']][:72]:72s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")

if __name__ == "__main__":
    (raw if sys.argv[1] == "raw" else source)(sys.argv[2])
