"""`write_data` of PyPIC3D/diagnostics/plotting.py:258-278: the text format of `data/total_energy.txt` and friends."""


def write_data(filename, time, data):
    """Append one `"<time>, <value>"` row (both printed as Python floats, like the reference's f-string)."""
    with open(filename, "a") as f:
        f.write(f"{float(time)}, {float(data)}\n")
