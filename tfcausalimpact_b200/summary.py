"""Text rendering of the causal-impact summary (presentation only; host side).

`summary(..., output='summary')` reproduces the reference's table layout byte for byte
(/root/reference/causalimpact/summary/templates/summary, golden files under tests/golden/).
`output='report'` gives a short prose interpretation written for this package.
"""
import pandas as pd

from .misc import get_z_score

_PAD = 19


def _fmt(x, digits):
    return str(round(float(x), digits))


def _sd(lower, upper, z, digits):
    return round((max(upper, lower) - min(upper, lower)) / (2 * z), digits)


def _ci_label(alpha):
    return str((1 - alpha) * 100).rstrip('0').rstrip('.') + '% CI'


def _pad(n):
    return ' ' * (_PAD - n)


def _render_summary(s, alpha, z, p_value, digits):
    a, c = s['average'], s['cumulative']
    r = lambda v: round(float(v), digits)   # noqa: E731
    lines = ['Posterior Inference {Causal Impact}', '                          Average            Cumulative']
    act = str(r(a['actual']))
    lines.append(f'Actual                    {act}{_pad(len(act))}{r(c["actual"])}')
    pa, sda = str(r(a['predicted'])), str(_sd(a['predicted_lower'], a['predicted_upper'], z, digits))
    sdc = _sd(c['predicted_lower'], c['predicted_upper'], z, digits)
    lines.append(f'Prediction (s.d.)         {pa} ({sda}){_pad(len(pa) + 3 + len(sda))}{r(c["predicted"])} ({sdc})')
    lo, hi = str(r(a['predicted_lower'])), str(r(a['predicted_upper']))
    lines.append(f'{_ci_label(alpha)}                    [{lo}, {hi}]{_pad(4 + len(lo) + len(hi))}'
                 f'[{r(c["predicted_lower"])}, {r(c["predicted_upper"])}]')
    lines.append('')
    ea, sea = str(r(a['abs_effect'])), str(_sd(a['abs_effect_lower'], a['abs_effect_upper'], z, digits))
    sec = _sd(c['abs_effect_lower'], c['abs_effect_upper'], z, digits)
    lines.append(f'Absolute effect (s.d.)    {ea} ({sea}){_pad(3 + len(ea) + len(sea))}{r(c["abs_effect"])} ({sec})')
    el, eh = r(a['abs_effect_lower']), r(a['abs_effect_upper'])
    lines.append(f'{_ci_label(alpha)}                    {sorted([el, eh])}{_pad(4 + len(str(el)) + len(str(eh)))}'
                 f'{sorted([r(c["abs_effect_lower"]), r(c["abs_effect_upper"])])}')
    lines.append('')
    ra = str(round(float(a['rel_effect']) * 100, digits))
    rsa = str(round(100 * float(_sd(a['rel_effect_lower'], a['rel_effect_upper'], z, 4)), digits))
    rc = round(100 * float(c['rel_effect']), digits)
    rsc = round(100 * float(_sd(c['rel_effect_lower'], c['rel_effect_upper'], z, 4)), digits)
    lines.append(f'Relative effect (s.d.)    {ra}% ({rsa}%){_pad(5 + len(ra) + len(rsa))}{rc}% ({rsc}%)')
    al, ah = round(float(a['rel_effect_lower']) * 100, digits), round(float(a['rel_effect_upper']) * 100, digits)
    cl, ch = round(100 * float(c['rel_effect_lower']), digits), round(100 * float(c['rel_effect_upper']), digits)
    lines.append(f'{_ci_label(alpha)}                    [{min(al, ah)}%, {max(ah, al)}%]'
                 f'{_pad(6 + len(str(min(al, ah))) + len(str(max(ah, al))))}[{min(cl, ch)}%, {max(ch, cl)}%]')
    lines.append('')
    lines.append(f'Posterior tail-area probability p: {round(float(p_value), digits)}')
    lines.append(f'Posterior prob. of a causal effect: {round((1 - float(p_value)) * 100, digits)}%')
    lines.append('')
    lines.append("For more details run the command: print(impact.summary('report'))")
    return '\n'.join(lines)


def _render_report(s, alpha, p_value, digits):
    a, c = s['average'], s['cumulative']
    r = lambda v: round(float(v), digits)   # noqa: E731
    ci = int(round((1 - alpha) * 100))
    sig = p_value < alpha
    direction = 'increase' if a['rel_effect'] > 0 else 'decrease'
    out = [
        'Analysis report {CausalImpact}', '',
        f'During the post-intervention period the response averaged {r(a["actual"])}; without the '
        f'intervention the model expected {r(a["predicted"])} ({ci}% interval [{r(a["predicted_lower"])}, '
        f'{r(a["predicted_upper"])}]). The estimated effect on the average is {r(a["abs_effect"])} '
        f'({ci}% interval [{r(a["abs_effect_lower"])}, {r(a["abs_effect_upper"])}]).', '',
        f'Summed over the post-intervention period the response was {r(c["actual"])} against an expected '
        f'{r(c["predicted"])} ({ci}% interval [{r(c["predicted_lower"])}, {r(c["predicted_upper"])}]).', '',
        f'In relative terms this is a {direction} of {r(a["rel_effect"] * 100)}% ({ci}% interval '
        f'[{r(a["rel_effect_lower"] * 100)}%, {r(a["rel_effect_upper"] * 100)}%]).', '',
        f'The posterior tail-area probability is p = {round(float(p_value), max(digits, 3))}: the effect is '
        + ('statistically significant and unlikely to be due to random fluctuation.' if sig else
           'not statistically significant and may be the result of random fluctuation.'),
    ]
    return '\n'.join(out)


def summary(summary_data: pd.DataFrame, p_value: float, alpha: float = 0.05, output: str = 'summary',
            digits: int = 2) -> str:
    if output not in {'summary', 'report'}:
        raise ValueError('Please choose either summary or report for output.')
    s = summary_data.to_dict()
    if output == 'summary':
        return _render_summary(s, alpha, get_z_score(1 - alpha / 2.), p_value, digits)
    return _render_report(s, alpha, p_value, digits)
