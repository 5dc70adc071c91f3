"""profiles/ from one `ncu --set full` capture of the four render kernels (scripts/gpu_prof.sh):
    python scripts/make_profiles.py gpurun_out/prof_r2b.ncu-rep r02 2
writes profiles/<tag>_render_kernels_full.txt (key metrics + top stalls per kernel, per-view instruction counts) and
profiles/traffic.json (dram bytes per view of the two main kernels, with the kernel build string the capture belongs to).
Runs without a GPU (ncu -i)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rep, tag, views = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'launch__shared_mem_per_block_static', 'local_load_store_bytes' ]


def num(r, name):
    v = r[h.index(name)].replace(',', '')
    u = units[h.index(name)]
    f = float(v)
    return f * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


from ava256_b200 import lib  # noqa: E402
build = lib.LIB.mvp_build_config().decode()
out = ["ncu --set full --clock-control none --import-source on, scripts/prof_step.py %d 1024 667 16384 8 2 (ALPHA_MU=17 ALPHA_SIGMA=6), B200." % views,
       "capture: %s   kernel build: %s" % (os.path.basename(rep), build), ""]
traffic = {"shape": [1024, 667, 16384, 8], "views_in_capture": views, "source": "profiles/%s_render_kernels_full.txt (%s)" % (tag, os.path.basename(rep)),
           "kernel_build": build}
for r in rows[2:]:
    name = r[h.index('Kernel Name')]
    out.append("Kernel Name  %s" % name)
    for w in want:
        if w in h:
            out.append("%-72s %s %s" % (w, r[h.index(w)], units[h.index(w)]))
    inst = num(r, 'smsp__inst_executed.sum')
    out.append("%-72s %.1f M" % ("warp instructions per view", inst / views / 1e6))
    dram = num(r, 'dram__bytes_read.sum') + num(r, 'dram__bytes_write.sum')
    out.append("%-72s %.1f MB" % ("dram bytes (read + write) per view", dram / views / 1e6))
    st = []
    for i, nm in enumerate(h):
        if 'issue_stalled' in nm and nm.endswith('_per_warp_active.pct'):
            try:
                st.append((float(r[i].replace(',', '')), nm))
            except ValueError:
                pass
    for v, nm in sorted(st, reverse=True)[:6]:
        out.append("   stall %-60s %6.2f" % (nm.split('issue_stalled_')[1].replace('_per_warp_active.pct', ''), v))
    out.append("")
    if ", 256," in name or "256, 0>" in name or "(int)256" in name:
        key = "render_forward_kernel" if "forward" in name else "render_backward_kernel"
        traffic[key] = {"dram_bytes_per_view": dram / views, "warp_instructions_per_view": inst / views}
        if key == "render_forward_kernel":
            # the captured forward runs in gradient mode: it also zero-fills the coming backward's gradient buffers (clear_grad_*:
            # K * (T^3 * 4 + 15) floats per view), bytes that used to be a separate memset pass and are not part of the
            # algorithmic traffic of SURVEY 8d
            clear = 16384 * (8 ** 3 * 4 + 15) * 4.0
            traffic[key]["of_which_gradient_zero_fill_per_view"] = clear
            traffic[key]["dram_bytes_per_view_without_zero_fill"] = dram / views - clear
open(os.path.join(ROOT, "profiles", "%s_render_kernels_full.txt" % tag), "w").write("\n".join(out))
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print("\n".join(out[:3]))
print(json.dumps(traffic)[:400])
