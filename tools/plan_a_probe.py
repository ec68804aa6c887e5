"""Plan-A probe: can the reference's own wgpu pipeline (crate
wgpu-3dgs-viewer 0.2.0 on a software Vulkan adapter) be built or run on THIS box?  Prints one JSON object.

    python tools/plan_a_probe.py [out.json]

Run here and on the gpurun box; bench.py embeds the result in `cpu_baseline.plan_a_probe`."""
import glob
import json
import os
import shutil
import socket
import subprocess
import sys


def _run(cmd, timeout=8):
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        return (r.stdout + r.stderr).strip()[:200]
    except Exception as e:  # noqa: BLE001
        return "failed: %s" % type(e).__name__


def probe():
    out = {}
    for tool in ("cargo", "rustc", "rustup", "trunk", "vulkaninfo", "glslangValidator", "naga", "curl"):
        out["which_" + tool] = shutil.which(tool)
    home = os.path.expanduser("~")
    reg = glob.glob(os.path.join(home, ".cargo", "registry", "*")) + glob.glob("/usr/local/cargo/registry/*")
    out["cargo_registry"] = reg[:4]
    crate = []
    for root in ("/root", "/usr", "/opt", "/home", "/var/cache"):
        if not os.path.isdir(root):
            continue
        r = _run(["find", root, "-xdev", "-maxdepth", "8", "(", "-name", "wgpu-3dgs-viewer*", "-o", "-name", "wgpu_3dgs_viewer*", ")",
                  "-not", "-path", "*/repo/*", "-not", "-path", "/root/reference/*"], timeout=60)
        crate += [x for x in r.splitlines() if x and "wgpu-3dgs-viewer-app_b200" not in x and not x.startswith("failed")]
    out["crate_source_found"] = crate[:8]
    icd = []
    for d in ("/usr/share/vulkan/icd.d", "/etc/vulkan/icd.d", "/usr/local/share/vulkan/icd.d"):
        icd += glob.glob(os.path.join(d, "*.json"))
    out["vulkan_icd"] = icd
    libs = []
    for pat in ("/usr/lib/x86_64-linux-gnu/libvulkan.so*", "/usr/lib/x86_64-linux-gnu/libvulkan_lvp*", "/usr/lib/x86_64-linux-gnu/libEGL.so*",
                "/usr/lib/x86_64-linux-gnu/libGLX_mesa*", "/usr/lib/x86_64-linux-gnu/dri/*swrast*"):
        libs += glob.glob(pat)
    out["software_adapter_libs"] = libs[:8]
    # network: 5-second TCP connect to crates.io (no data sent)
    try:
        s = socket.create_connection(("crates.io", 443), timeout=5)
        s.close()
        out["crates_io_reachable"] = True
    except Exception as e:  # noqa: BLE001
        out["crates_io_reachable"] = False
        out["crates_io_error"] = type(e).__name__
    out["baseline_ref_dir"] = os.path.isdir(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref"))
    out["nproc"] = os.cpu_count()
    out["runnable"] = bool(out["which_cargo"] and (out["cargo_registry"] or out["crates_io_reachable"]) and
                           (out["vulkan_icd"] or out["software_adapter_libs"]))
    out["hostname_kind"] = "gpu-box" if shutil.which("nvidia-smi") and "failed" not in _run(["nvidia-smi", "-L"]) and \
        "GPU" in _run(["nvidia-smi", "-L"]) else "cpu-container"
    return out


if __name__ == "__main__":
    res = probe()
    txt = json.dumps(res, indent=1)
    if len(sys.argv) > 1:
        os.makedirs(os.path.dirname(os.path.abspath(sys.argv[1])), exist_ok=True)
        open(sys.argv[1], "w").write(txt + "\n")
    print(txt)
