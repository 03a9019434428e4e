"""CPU-only hygiene checks of the measurement plumbing: the committed DRAM-traffic capture describes THIS tree's kernel sources (bench.py
prints roofline.traffic only then), and the ncu log helpers under tools/ parse what ncu writes."""
import json
import os
import subprocess
import sys

from conftest import ROOT

NCU_CSV = '''==PROF== Connected to process 1 (/usr/bin/python3.12)
"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"
"0","1","python3.12","127.0.0.1","void bm::frame_kernel_q<1, 0>(bm::FrameParams)","1","7","(1024, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_read.sum","Mbyte","46.7"
"0","1","python3.12","127.0.0.1","void bm::frame_kernel_q<1, 0>(bm::FrameParams)","1","7","(1024, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_write.sum","Mbyte","62.2"
"0","1","python3.12","127.0.0.1","void bm::frame_kernel_q<1, 0>(bm::FrameParams)","1","7","(1024, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","us","746.0"
"1","1","python3.12","127.0.0.1","bm::scan_kernel(bm::FrameIO)","1","7","(256, 1, 1)","(32, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","ns","5700"
'''


def test_traffic_capture_describes_this_trees_kernel_sources():
    """profiles/frame_kernel_traffic.json is stamped with a hash of brickmap_b200/csrc; a kernel change without a new capture must not
    go unnoticed (the traffic figure was a stale constant)."""
    from brickmap_b200.build import source_hash
    with open(os.path.join(ROOT, "profiles", "frame_kernel_traffic.json")) as f:
        tj = json.load(f)
    assert tj["source_hash"] == source_hash(), "re-capture the frame kernel (tools/gpu_r2_final.sh) after changing brickmap_b200/csrc"
    for name in ("cfg3", "cfg4"):
        e = tj["configs"][name]
        l0 = e["launches"][0]
        assert "frame_kernel_q" in l0["kernel"]
        assert e["dram_bytes_per_launch"] == l0["dram_bytes_read"] + l0["dram_bytes_write"] > 0
        assert e["steady"]["launches"] >= 3 and e["steady"]["dram_bytes_per_launch"] > 0
        assert len(e["steady"]["dram_bytes_of_each_launch"]) == e["steady"]["launches"]


def test_ncu_log_helpers_parse_an_ncu_csv(tmp_path):
    log = tmp_path / "launches.csv"
    log.write_text(NCU_CSV)
    rows = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_csv_rows.py"), str(log)], capture_output=True, text=True, check=True).stdout
    assert rows.startswith("46.7 R + 62.2 W MB, 746 us | ")
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), str(log)], capture_output=True, text=True, check=True).stdout
    assert "frame_kernel_q : scan_kernel = 130.9 : 1" in summ
    assert "launches=   1  total      746.0 us" in summ
