"""Known answers printed by R itself for lm() / glm(binomial) on the `mtcars` data set that ships with base R
(datasets::mtcars; the summary() outputs below are reproduced in countless R tutorials and are stable across R versions).
They pin the restatement of lm.fit / glm.fit / summary.* in oracle/npdr_oracle.c -- and, through it, the GPU solvers -- to
outputs of the real reference stack, including the number of Fisher scoring iterations (i.e. glm.fit's start values and
deviance stopping rule).  Values are given to the digits R prints.

    summary(lm(mpg ~ wt, mtcars))                 coef 37.2851 / -5.3445, se 1.8776 / 0.5591, t 19.858 / -9.559, R^2 0.7528
    summary(lm(mpg ~ hp + wt, mtcars))            coef 37.22727 / -0.03177 / -3.87783, se 1.59879 / 0.00903 / 0.63273, R^2 0.8268
    summary(glm(am ~ wt, binomial, mtcars))       coef 12.040 / -4.024, se 4.510 / 1.436, z 2.670 / -2.801,
                                                  residual deviance 19.176 on 30 df, Fisher scoring iterations: 6
    summary(glm(am ~ hp + wt, binomial, mtcars))  coef 18.86630 / 0.03626 / -8.08348, se 7.44356 / 0.01773 / 3.06868,
                                                  residual deviance 10.059 on 29 df, Fisher scoring iterations: 8
"""
import numpy as np

MPG = np.array([21.0, 21.0, 22.8, 21.4, 18.7, 18.1, 14.3, 24.4, 22.8, 19.2, 17.8, 16.4, 17.3, 15.2, 10.4, 10.4, 14.7, 32.4, 30.4,
                33.9, 21.5, 15.5, 15.2, 13.3, 19.2, 27.3, 26.0, 30.4, 15.8, 19.7, 15.0, 21.4])
WT = np.array([2.620, 2.875, 2.320, 3.215, 3.440, 3.460, 3.570, 3.190, 3.150, 3.440, 3.440, 4.070, 3.730, 3.780, 5.250, 5.424,
               5.345, 2.200, 1.615, 1.835, 2.465, 3.520, 3.435, 3.840, 3.845, 1.935, 2.140, 1.513, 3.170, 2.770, 3.570, 2.780])
HP = np.array([110, 110, 93, 110, 175, 105, 245, 62, 95, 123, 123, 180, 180, 180, 205, 215, 230, 66, 52, 65, 97, 150, 150, 245,
               175, 66, 91, 113, 264, 175, 335, 109], dtype=float)
AM = np.array([1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1], dtype=float)

# name -> (response, predictor columns in formula order, family, printed coefficients, printed std. errors, extras)
KATS = {
    "lm_mpg_wt": (MPG, [WT], "lm", [37.2851, -5.3445], [1.8776, 0.5591], {"r_squared": 0.7528, "df": 30}),
    "lm_mpg_hp_wt": (MPG, [HP, WT], "lm", [37.22727, -0.03177, -3.87783], [1.59879, 0.00903, 0.63273], {"r_squared": 0.8268, "df": 29}),
    "glm_am_wt": (AM, [WT], "binomial", [12.040, -4.024], [4.510, 1.436], {"deviance": 19.176, "df": 30, "iters": 6}),
    "glm_am_hp_wt": (AM, [HP, WT], "binomial", [18.86630, 0.03626, -8.08348], [7.44356, 0.01773, 3.06868],
                     {"deviance": 10.059, "df": 29, "iters": 8}),
}


def printed_close(value, printed):
    """True if `value` rounds to the number R printed (same count of decimals)."""
    s = repr(printed)
    decimals = len(s.split(".")[1]) if "." in s else 0
    return abs(value - printed) <= 0.5000001 * 10.0 ** (-decimals)
