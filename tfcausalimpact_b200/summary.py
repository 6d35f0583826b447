"""Text rendering of the causal-impact summary (presentation only; host side).

`summary(..., output='summary')` reproduces the reference's table layout byte for byte
(/root/reference/causalimpact/summary/templates/summary, golden files under tests/golden/).
`output='report'` reproduces the reference's report wording (templates/report), pinned by the five
report goldens of the reference's test-suite.
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
    """The reference's prose report (wording, paragraph variants and rounding of
    /root/reference/causalimpact/summary/templates/report:1-81; pinned byte for byte by the five
    tests/golden/test_report_summary_* files of the reference's test-suite).
    Significance follows the reference: the (1 - alpha) interval of the relative effect excludes 0
    (`detected_sig`), sign of the relative effect (`positive_sig`); the p-value only selects the closing
    paragraph."""
    a, c = s['average'], s['cumulative']
    r = lambda v: round(float(v), digits)   # noqa: E731
    detected = not (a['rel_effect_lower'] < 0 and a['rel_effect_upper'] > 0)
    positive = a['rel_effect'] > 0
    ci = str((1 - alpha) * 100).rstrip('0').rstrip('.') + '%'
    rel = [r(100 * a['rel_effect_lower']), r(100 * a['rel_effect_upper'])]
    p = [
        'Analysis report {CausalImpact}\n\n\n'
        'During the post-intervention period, the response variable had\n'
        f'an average value of approx. {r(a["actual"])}. {"By contrast, in" if detected else "In"} the absence of an\n'
        f'intervention, we would have expected an average response of {r(a["predicted"])}.\n'
        f'The {ci} interval of this counterfactual prediction is [{r(a["predicted_lower"])}, {r(a["predicted_upper"])}].\n'
        'Subtracting this prediction from the observed response yields\n'
        'an estimate of the causal effect the intervention had on the\n'
        f'response variable. This effect is {r(a["abs_effect"])} with a {ci} interval of\n'
        f'{sorted([r(a["abs_effect_lower"]), r(a["abs_effect_upper"])])}. For a discussion of the significance of this effect,\n'
        'see below.\n\n\n'
        'Summing up the individual data points during the post-intervention\n'
        'period (which can only sometimes be meaningfully interpreted), the\n'
        f'response variable had an overall value of {r(c["actual"])}.\n'
        f'{"By contrast, had" if detected else "Had"} the intervention not taken place, we would have expected\n'
        f'a sum of {r(c["predicted"])}. The {ci} interval of this prediction is '
        f'{sorted([r(c["predicted_lower"]), r(c["predicted_upper"])])}.\n\n\n'
        'The above results are given in terms of absolute numbers. In relative\n'
        f'terms, the response variable showed {"an increase of +" if positive else "a decrease of "}'
        f'{r(100 * a["rel_effect"])}%. The {ci}\n'
        f'interval of this percentage is [{min(rel)}%, {max(rel)}%].\n'
    ]
    if detected and positive:
        p.append('\n\nThis means that the positive effect observed during the intervention\n'
                 'period is statistically significant and unlikely to be due to random\n'
                 'fluctuations. It should be noted, however, that the question of whether\n'
                 'this increase also bears substantive significance can only be answered\n'
                 f'by comparing the absolute effect ({r(a["abs_effect"])}) to the original goal\n'
                 'of the underlying intervention.\n')
    elif detected:
        p.append('\n\nThis means that the negative effect observed during the intervention\n'
                 'period is statistically significant.\n'
                 'If the experimenter had expected a positive effect, it is recommended\n'
                 'to double-check whether anomalies in the control variables may have\n'
                 'caused an overly optimistic expectation of what should have happened\n'
                 'in the response variable in the absence of the intervention.\n')
    elif positive:
        p.append('\n\nThis means that, although the intervention appears to have caused a\n'
                 'positive effect, this effect is not statistically significant when\n'
                 'considering the entire post-intervention period as a whole. Individual\n'
                 'days or shorter stretches within the intervention period may of course\n'
                 'still have had a significant effect, as indicated whenever the lower\n'
                 'limit of the impact time series (lower plot) was above zero.\n')
    else:
        p.append('This means that, although it may look as though the intervention has\n'
                 'exerted a negative effect on the response variable when considering\n'
                 'the intervention period as a whole, this effect is not statistically\n'
                 'significant and so cannot be meaningfully interpreted.\n')
    if not detected:
        p.append('\n\nThe apparent effect could be the result of random fluctuations that\n'
                 'are unrelated to the intervention. This is often the case when the\n'
                 'intervention period is very long and includes much of the time when\n'
                 'the effect has already worn off. It can also be the case when the\n'
                 'intervention period is too short to distinguish the signal from the\n'
                 'noise. Finally, failing to find a significant effect can happen when\n'
                 'there are not enough control variables or when these variables do not\n'
                 'correlate well with the response variable during the learning period.\n')
    if p_value < alpha:
        p.append('\n\nThe probability of obtaining this effect by chance is very small\n'
                 f'(Bayesian one-sided tail-area probability p = {round(float(p_value), digits)}).\n'
                 'This means the causal effect can be considered statistically\n'
                 'significant.')
    else:
        p.append(f'\n\nThe probability of obtaining this effect by chance is p = {round(float(p_value) * 100, digits)}%.\n'
                 'This means the effect may be spurious and would generally not be\n'
                 'considered statistically significant.')
    return ''.join(p)


def summary(summary_data: pd.DataFrame, p_value: float, alpha: float = 0.05, output: str = 'summary',
            digits: int = 2) -> str:
    if output not in {'summary', 'report'}:
        raise ValueError('Please choose either summary or report for output.')
    s = summary_data.to_dict()
    if output == 'summary':
        return _render_summary(s, alpha, get_z_score(1 - alpha / 2.), p_value, digits)
    return _render_report(s, alpha, p_value, digits)
