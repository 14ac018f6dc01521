"""Convert the reference's bundled datasets to `.npz` fixtures (run once, in the
build container; needs /root/reference).  Test infrastructure only.

    python tests/golden/make_fixtures.py

Sources (documented in /root/reference/R/data.R:1-53, produced by
/root/reference/data-raw/fetch-data.R:10-43):
  case.control.3sets$train  -> case_control_train.npz   (BASELINE config 1)
  qtrait.3sets$train        -> qtrait_train.npz         (BASELINE config 2)
  mdd.RNAseq.small          -> mdd_rnaseq_small.npz     (BASELINE config 3)
  testdata10                -> testdata10.npz           (tiny edge-case set)
Matrices are stored column-major (Fortran order) float64 exactly as R holds them.
Factors are stored as their 1-based integer codes plus the level strings.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from rda_reader import as_factor, data_frame_columns, names, read_rda  # noqa: E402

REF = "/root/reference/data"


def df_to_matrix(df, drop):
    cols = data_frame_columns(df)
    keep = [c for c in cols if c not in drop]
    X = np.asfortranarray(np.column_stack([cols[c].value for c in keep]).astype(np.float64))
    return X, keep, cols


def main():
    # config 1: case/control
    cc = read_rda(f"{REF}/case.control.3sets.rda")["case.control.3sets"]
    parts = dict(zip(names(cc), cc.value))
    X, attr_names, cols = df_to_matrix(parts["train"], drop=("class",))
    codes, levels = as_factor(cols["class"])
    np.savez_compressed(f"{HERE}/case_control_train.npz", X=X, attr_names=np.array(attr_names),
                        class_codes=codes, class_levels=np.array(levels),
                        signal_names=np.array(list(parts["signal.names"].value)))
    print("case_control_train", X.shape, levels, np.bincount(codes)[1:])

    # config 2: quantitative trait
    qt = read_rda(f"{REF}/qtrait.3sets.rda")["qtrait.3sets"]
    parts = dict(zip(names(qt), qt.value))
    X, attr_names, cols = df_to_matrix(parts["train"], drop=("qtrait",))
    np.savez_compressed(f"{HERE}/qtrait_train.npz", X=X, attr_names=np.array(attr_names),
                        qtrait=cols["qtrait"].value.astype(np.float64),
                        signal_names=np.array(list(parts["signal.names"].value)))
    print("qtrait_train", X.shape)

    # config 3: RNA-seq with covariates
    mdd = read_rda(f"{REF}/mdd.RNAseq.small.rda")["mdd.RNAseq.small"]
    parts = dict(zip(names(mdd), mdd.value))
    rna = parts["rnaSeq"]
    dim = tuple(int(v) for v in rna.attr["dim"].value)
    X = np.asfortranarray(rna.value.reshape(dim, order="F"))
    gene_names = list(rna.attr["dimnames"].value[1].value)
    ph_codes, ph_levels = as_factor(parts["phenos"])
    covs = data_frame_columns(parts["covs.short"])
    sex_codes, sex_levels = as_factor(covs["sex"])
    age = covs["age"].value.astype(np.float64)
    np.savez_compressed(f"{HERE}/mdd_rnaseq_small.npz", X=X, attr_names=np.array(gene_names),
                        pheno_codes=ph_codes, pheno_levels=np.array(ph_levels),
                        sex_codes=sex_codes, sex_levels=np.array(sex_levels), age=age)
    print("mdd", X.shape, ph_levels, np.bincount(ph_codes)[1:], sex_levels, age[:4])

    # tiny set
    td = read_rda(f"{REF}/testdata10.rda")["testdata10"]
    X, attr_names, cols = df_to_matrix(td, drop=("Class",))
    np.savez_compressed(f"{HERE}/testdata10.npz", X=X, attr_names=np.array(attr_names),
                        Class=np.asarray(cols["Class"].value, dtype=np.float64))
    print("testdata10", X.shape, cols["Class"].value)


if __name__ == "__main__":
    main()
