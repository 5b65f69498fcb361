"""Solid masks for the volume penalisation and running statistics (host side).

`navier.solid = solid_cylinder_inner(navier.temp.x[0], navier.temp.x[1], 0.2, 0.0, 0.2)` like the reference
(src/navier/solid_masks.rs:34-175); `navier.statistics = Statistics(navier, save_stat, write_stat)`
(src/navier/statistics.rs:10-247).  The masks are plain [nx, ny] arrays on the physical grid; the penalisation itself
runs in the fused product kernel on the device (rp_navier_set_solid).
"""
import math

import numpy as np

from . import snapshot


def solid_cylinder_inner(x, y, x0, y0, radius):
    """Everything with r < radius is solid; tanh smoothing layer of radius / 10 (solid_masks.rs:34-57)."""
    x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
    r = np.sqrt((x0 - x)[:, None] ** 2.0 + (y0 - y)[None, :] ** 2.0)
    layer = radius / 10.0
    mask = np.where(r < radius - layer, 1.0, np.where(r < radius + layer, 0.5 * (1.0 - np.tanh(2.0 * (r - radius) / layer)), 0.0))
    return [mask, np.zeros_like(mask)]


def solid_roughness_sinusoid(x, y, height, wavenumber):
    """Sinusoidal roughness elements on both plates, value = +-0.5 inside (solid_masks.rs:59-98)."""
    x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
    mask = np.zeros((len(x), len(y)))
    value = np.zeros((len(x), len(y)))
    bottom, top = y[0], y[-1]
    layer = height / 10.0
    y_rough = (height * (top - bottom) / 2.0 * (np.sin(wavenumber * x) + 0.5))[:, None]
    for y_dist, val in (((y - bottom)[None, :], 0.5), ((top - y)[None, :], -0.5)):  # bottom first, then top (overrides)
        inside = y_dist <= y_rough
        smooth = (~inside) & (y_dist <= y_rough + layer)
        mask = np.where(inside, 1.0, np.where(smooth, 0.5 * (1.0 - np.tanh(2.0 * (y_dist - y_rough) / layer)), mask))
        value = np.where(inside | smooth, val, value)
    return [mask, value]


def _round_half_away(v):
    return math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5)


def solid_porosity(x, y, diameter, porosity):
    """Regular array of circles that mimics a porous medium (solid_masks.rs:100-136)."""
    x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
    mask = np.zeros((len(x), len(y)))
    radius = diameter / 2.0
    length, height = x[-1] - x[0], y[-1] - y[0]
    ncx = _round_half_away(math.sqrt((1.0 - porosity) * 4.0 * length ** 2 / (math.pi * diameter ** 2)))
    ncy = _round_half_away(math.sqrt((1.0 - porosity) * 4.0 * height ** 2 / (math.pi * diameter ** 2)))
    dx = (length - ncx * diameter) / (ncx + 1.0)
    dy = (height - ncy * diameter) / (ncy + 1.0)
    ox = x[0] + dx + radius
    for _ in range(int(ncx)):
        oy = y[0] + dy + radius
        for _ in range(int(ncy)):
            mask += solid_cylinder_inner(x, y, ox, oy, radius)[0]
            oy += dy + diameter
        ox += dx + diameter
    return [mask, np.zeros_like(mask)]


def solid_porosity_interpolate(nx, ny, diameter, porosity, lib=None):
    """Porosity mask built on a 513 x 513 Chebyshev grid and spectrally interpolated (solid_masks.rs:138-162)."""
    from . import api
    src = api.Field2(api.Space2(api.chebyshev(513), api.chebyshev(513)), lib=lib)
    dst = api.Field2(api.Space2(api.chebyshev(nx), api.chebyshev(ny)), lib=lib)
    x, y = src.x
    out = []
    for a in solid_porosity(x, y, diameter, porosity):
        src.v = a
        src.forward()
        old, new = src.vhat, np.array(dst.vhat)
        r, c = min(old.shape[0], new.shape[0]), min(old.shape[1], new.shape[1])
        new[:r, :c] = old[:r, :c]  # broadcast_2d (solid_masks.rs:165-175)
        dst.vhat = new
        dst.backward()
        out.append(dst.v)
    return out


class Statistics:
    """statistics.rs:10-160: running average of T, the last ux / uy, the Nusselt field -- ortho coefficients on the
    space of `navier.field`; the transforms run on the device through Field2."""

    def __init__(self, navier, save_stat, write_stat):
        self.nu, self.ka, self.ra, self.pr = navier.nu, navier.ka, navier.ra, navier.pr
        self.scale = list(navier.scale)
        mk = navier.new_work_field
        self.field, self.t_avg, self.ux_avg, self.uy_avg, self.nusselt = mk(), mk(), mk(), mk(), mk()
        self.save_stat, self.write_stat = save_stat, write_stat
        self.avg_time = 0.0
        self.tot_time = navier.time
        self.num_save = 0

    def update(self, that, uxhat, uyhat, time):  # statistics.rs:130-159
        if time < self.tot_time:
            print("Statistics time mismatch (navier < stat): %r < %r" % (time, self.tot_time))
            return
        weight = float(self.num_save)
        self.t_avg.vhat = (self.t_avg.vhat * weight + that) / (weight + 1.0)
        self.ux_avg.vhat = uxhat
        self.uy_avg.vhat = uyhat
        f = self.field  # nusselt(), statistics.rs:215-247
        f.vhat = uyhat
        f.backward()
        uy_v = f.v
        f.vhat = that
        f.backward()
        uy_temp = f.v * uy_v
        f.vhat = f.gradient([0, 1], None) / (self.scale[1] * -1.0)
        f.backward()
        f.v = (f.v + uy_temp / self.ka) * 2.0 * self.scale[1]
        f.forward()
        self.nusselt.vhat = f.vhat
        self.num_save += 1
        self.avg_time += time - self.tot_time
        self.tot_time = time

    def write(self, filename):  # statistics.rs:163-194 (RPSNAP1 container, same dataset names)
        out = {}
        for g, f in (("temp", self.t_avg), ("ux", self.ux_avg), ("uy", self.uy_avg), ("nusselt", self.nusselt)):
            f.backward()
            out[g + "/v"] = f.v
            vh = f.vhat
            if np.iscomplexobj(vh):
                out[g + "/vhat_re"], out[g + "/vhat_im"] = vh.real.copy(), vh.imag.copy()
            else:
                out[g + "/vhat"] = vh
            x, dx = f.x, f.dx
            out["x"], out["dx"], out["y"], out["dy"] = x[0], dx[0], x[1], dx[1]
        out.update(tot_time=self.tot_time, avg_time=self.avg_time, num_save=float(self.num_save), ra=self.ra, pr=self.pr, nu=self.nu, ka=self.ka)
        snapshot.write_snapshot(filename, out)
        print(" ==> %r" % str(filename))

    def read(self, filename):  # statistics.rs:197-211
        d = snapshot.read_snapshot(filename)
        for g, f in (("temp", self.t_avg), ("ux", self.ux_avg), ("uy", self.uy_avg), ("nusselt", self.nusselt)):
            vh = d[g + "/vhat_re"] + 1j * d[g + "/vhat_im"] if (g + "/vhat_re") in d else d[g + "/vhat"]
            new = np.array(f.vhat)
            r, c = min(vh.shape[0], new.shape[0]), min(vh.shape[1], new.shape[1])
            new[:r, :c] = vh[:r, :c]
            f.vhat = new
            f.backward()
        self.tot_time, self.avg_time, self.num_save = float(d["tot_time"]), float(d["avg_time"]), int(d["num_save"])
        print(" <== %r" % str(filename))
