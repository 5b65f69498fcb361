"""Solid masks and statistics -- TEST INFRASTRUCTURE (CPU restatement of the reference, see oracle/__init__.py).

solid masks: src/navier/solid_masks.rs:34-175; statistics: src/navier/statistics.rs:10-247."""
import math

import numpy as np


def solid_cylinder_inner(x, y, x0, y0, radius):  # solid_masks.rs:34-57
    mask = np.zeros((len(x), len(y)))
    layer = radius / 10.0
    for i, xi in enumerate(x):
        for j, yi in enumerate(y):
            r = math.sqrt((x0 - xi) ** 2.0 + (y0 - yi) ** 2.0)
            if r < radius - layer:
                mask[i, j] = 1.0
            elif r < radius + layer:
                mask[i, j] = 0.5 * (1.0 - math.tanh(2.0 * (r - radius) / layer))
    return [mask, np.zeros_like(mask)]


def solid_roughness_sinusoid(x, y, height, wavenumber):  # solid_masks.rs:59-98
    mask = np.zeros((len(x), len(y)))
    value = np.zeros((len(x), len(y)))
    bottom, top = y[0], y[-1]
    layer = height / 10.0
    for i, xi in enumerate(x):
        y_rough = height * (top - bottom) / 2.0 * (math.sin(wavenumber * xi) + 0.5)
        for j, yi in enumerate(y):
            for y_dist, val in ((yi - bottom, 0.5), (top - yi, -0.5)):
                if y_dist <= y_rough:
                    mask[i, j] = 1.0
                    value[i, j] = val
                elif y_dist <= y_rough + layer:
                    mask[i, j] = 0.5 * (1.0 - math.tanh(2.0 * (y_dist - y_rough) / layer))
                    value[i, j] = val
    return [mask, value]


def solid_porosity(x, y, diameter, porosity):  # solid_masks.rs:100-136
    mask = np.zeros((len(x), len(y)))
    radius = diameter / 2.0
    length, height = x[-1] - x[0], y[-1] - y[0]

    def rround(v):  # f64::round: half away from zero
        return math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5)

    ncx = rround(math.sqrt((1.0 - porosity) * 4.0 * length ** 2 / (math.pi * diameter ** 2)))
    ncy = rround(math.sqrt((1.0 - porosity) * 4.0 * height ** 2 / (math.pi * diameter ** 2)))
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


class Statistics:
    """statistics.rs:10-160: running average of T, last ux / uy, Nusselt field; all in ortho coefficients of `field`."""

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
        self.ux_avg.vhat = np.array(uxhat)
        self.uy_avg.vhat = np.array(uyhat)
        self._nusselt(that, uyhat)
        self.nusselt.vhat = np.array(self.field.vhat)
        self.num_save += 1
        self.avg_time += time - self.tot_time
        self.tot_time = time

    def _nusselt(self, that, uyhat):  # statistics.rs:215-247
        f = self.field
        f.vhat = np.array(uyhat)
        f.backward()
        uy_v = np.array(f.v)
        f.vhat = np.array(that)
        f.backward()
        uy_temp = np.array(f.v) * uy_v
        dtdz = f.gradient([0, 1], None) / (self.scale[1] * -1.0)
        f.vhat = dtdz
        f.backward()
        f.v = (np.array(f.v) + uy_temp / self.ka) * 2.0 * self.scale[1]
        f.forward()
