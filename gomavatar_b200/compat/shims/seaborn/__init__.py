"""Stand-in for ``seaborn`` (absent offline): ``color_palette`` for the three palettes the reference asks for
(eval.py:28 "hls", models/model.py:262 "tab10", utils/tb_util.py:142 "coolwarm") — visualisation colours only."""
import colorsys

_TAB10 = [(0.12156862745098039, 0.4666666666666667, 0.7058823529411765), (1.0, 0.4980392156862745, 0.054901960784313725),
          (0.17254901960784313, 0.6274509803921569, 0.17254901960784313), (0.8392156862745098, 0.15294117647058825, 0.1568627450980392),
          (0.5803921568627451, 0.403921568627451, 0.7411764705882353), (0.5490196078431373, 0.33725490196078434, 0.29411764705882354),
          (0.8901960784313725, 0.4666666666666667, 0.7607843137254902), (0.4980392156862745, 0.4980392156862745, 0.4980392156862745),
          (0.7372549019607844, 0.7411764705882353, 0.13333333333333333), (0.09019607843137255, 0.7450980392156863, 0.8117647058823529)]


def hls_palette(n_colors=6, h=.01, l=.6, s=.65):
    hues = [((i / n_colors) + h) % 1.0 for i in range(n_colors)]
    return [colorsys.hls_to_rgb(hh, l, s) for hh in hues]


def _coolwarm(n):
    cold, mid, warm = (0.2298, 0.2987, 0.7537), (0.8650, 0.8650, 0.8650), (0.7057, 0.0156, 0.1502)
    out = []
    for i in range(n):
        t = (i + 0.5) / n
        a, b, u = (cold, mid, t * 2) if t < 0.5 else (mid, warm, t * 2 - 1)
        out.append(tuple(a[c] + (b[c] - a[c]) * u for c in range(3)))
    return out


def color_palette(palette=None, n_colors=None, desat=None, as_cmap=False):
    if palette is None or palette == "tab10":
        n = 10 if n_colors is None else n_colors
        return [_TAB10[i % 10] for i in range(n)]
    if palette == "hls":
        return hls_palette(6 if n_colors is None else n_colors)
    if palette == "coolwarm":
        return _coolwarm(6 if n_colors is None else n_colors)
    raise NotImplementedError(f"seaborn stand-in: palette {palette!r} (install seaborn)")
