"""Development tool (runs where ncu is installed, no GPU needed): turn ncu reports brought back from the GPU box into the small JSON
bench.py reads for `roofline.traffic` and the tracking issue-slot roofline.

    python scripts/ncu_extract.py <infer.ncu-rep> [<gen_rays.ncu-rep>] > profiles/rNN_dominant_kernel_ncu.json
"""
import csv, io, json, subprocess, sys

N_INFER = 1920 * 1080


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]]


def f(row, key):
    return float(row[key].replace(",", ""))


def unit_bytes(row, key, hdr_units):
    return f(row, key)


def main():
    rows = raw(sys.argv[1])
    units = None
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rr = list(csv.reader(io.StringIO(out)))
    units = dict(zip(rr[0], rr[1]))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    r = rows[-1]
    rd = f(r, "dram__bytes_read.sum") * scale[units["dram__bytes_read.sum"]]
    wr = f(r, "dram__bytes_write.sum") * scale[units["dram__bytes_write.sum"]]
    j = {"kernel": r["Kernel Name"], "source": sys.argv[1].split("/")[-1], "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
         "duration_us_under_ncu": f(r, "gpu__time_duration.sum"),
         "l2": {"sectors_per_query": f(r, "lts__t_sectors_srcunit_tex_op_read.sum") / N_INFER,
                "l1_to_l2_request_unit_busy": f(r, "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed") / 100,
                "lts_throughput": f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed") / 100,
                "lts_hit_rate": f(r, "lts__t_sector_hit_rate.pct") / 100, "l1_hit_rate": f(r, "l1tex__t_sector_hit_rate.pct") / 100},
         "issue_slots_busy": f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") / 100,
         "alu_pipe": f(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active") / 100,
         "tensor_pipe": f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") / 100,
         "warps_active": f(r, "sm__warps_active.avg.pct_of_peak_sustained_active") / 100,
         "registers_per_thread": int(f(r, "launch__registers_per_thread"))}
    if len(sys.argv) > 2:
        g = raw(sys.argv[2])[-1]
        j["gen_rays"] = {"kernel": g["Kernel Name"], "source": sys.argv[2].split("/")[-1], "bound": "instruction issue", "duration_us_under_ncu": f(g, "gpu__time_duration.sum"),
                         "issue_slots_busy": f(g, "smsp__issue_active.avg.pct_of_peak_sustained_active") / 100,
                         "active_threads_per_warp": f(g, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                         "lts_throughput": f(g, "lts__throughput.avg.pct_of_peak_sustained_elapsed") / 100,
                         "dram_throughput": f(g, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") / 100}
    print(json.dumps(j, indent=1))


if __name__ == "__main__":
    main()
