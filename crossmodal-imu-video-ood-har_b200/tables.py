"""OOD results tables: the piece the reference's ``generate_tables.py`` lacks.

``generate_tables.TableGenerator`` (reference generate_tables.py:12-199) pivots few-shot result CSVs into
"mean +/- std" tables and writes them as CSV / LaTeX / Markdown; it has no OOD table because the reference
computes no OOD metric.  ``generate_ood_table`` produces that table in the same style from the
rows ``Evaluator.evaluate_ood`` returns, and ``save_tables`` writes the same three formats with the same file
naming (``<prefix>_<name>.{csv,tex,md}``), so a maintainer can call it next to ``generate_table3_style`` inside
``create_article_tables_from_results`` (reference generate_tables.py:202-335) -- see INTEGRATION.md.

Expected columns of ``results_df`` (one row per scorer x split x run):
    experiment  e.g. "cross_modal_pretrained"
    split       name of the ID-vs-held-out-activity split, e.g. "holdout_8_of_32"
    scorer      "msp" | "energy" | "maha"
    run         run index (optional)
    auroc       0..1
    fpr95       0..1   (FPR at the first threshold with TPR >= 0.95)
Host-side pandas only; nothing here touches a GPU.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Iterable, List, Mapping, Union

import pandas as pd

__all__ = ["format_mean_std", "ood_rows", "generate_ood_table", "save_tables"]


def format_mean_std(mean: float, std: float, decimals: int = 2) -> str:
    """"mean +/- std" in the reference's format (generate_tables.py:22-26); a single run prints the mean alone."""
    if std != std:            # NaN: one run only
        return f"{mean:.{decimals}f}"
    return f"{mean:.{decimals}f} ± {std:.{decimals}f}"


def ood_rows(experiment: str, split: str, results: Mapping[str, Mapping[str, float]], run: int = 0) -> List[dict]:
    """Flattens ``Evaluator.evaluate_ood`` output ({scorer: {'auroc','fpr', ...}}) into table rows."""
    rows = []
    for scorer, r in results.items():
        rows.append({"experiment": experiment, "split": split, "scorer": scorer, "run": run,
                     "auroc": float(r["auroc"]), "fpr95": float(r.get("fpr95", r.get("fpr")))})
    return rows


def generate_ood_table(results_df: pd.DataFrame, decimals: int = 2) -> Dict[str, pd.DataFrame]:
    """Returns {'ood_auroc', 'ood_fpr95', 'ood_summary'}: the first two are experiment x (split, scorer) pivots
    of "mean +/- std" percentages; the summary is the long table with both metrics side by side."""
    need = {"experiment", "split", "scorer", "auroc", "fpr95"}
    missing = need - set(results_df.columns)
    if missing:
        raise ValueError(f"generate_ood_table: missing columns {sorted(missing)}")
    g = results_df.groupby(["experiment", "split", "scorer"], sort=True).agg(
        auroc_mean=("auroc", "mean"), auroc_std=("auroc", "std"), fpr95_mean=("fpr95", "mean"),
        fpr95_std=("fpr95", "std"), runs=("auroc", "size")).reset_index()
    g["AUROC (%)"] = [format_mean_std(100 * m, 100 * s, decimals) for m, s in zip(g.auroc_mean, g.auroc_std)]
    g["FPR95 (%)"] = [format_mean_std(100 * m, 100 * s, decimals) for m, s in zip(g.fpr95_mean, g.fpr95_std)]
    tables = {
        "ood_auroc": g.pivot(index="experiment", columns=["split", "scorer"], values="AUROC (%)"),
        "ood_fpr95": g.pivot(index="experiment", columns=["split", "scorer"], values="FPR95 (%)"),
        "ood_summary": g[["experiment", "split", "scorer", "runs", "AUROC (%)", "FPR95 (%)"]].set_index(
            ["experiment", "split", "scorer"]),
    }
    return tables


def save_tables(tables: Mapping[str, pd.DataFrame], results_dir: Union[str, Path], prefix: str = "table") -> List[str]:
    """CSV + LaTeX + Markdown per table, named like the reference's ``TableGenerator.save_tables``
    (generate_tables.py:161-199).  Markdown falls back to a plain pipe table when ``tabulate`` is missing."""
    out = Path(results_dir)
    out.mkdir(parents=True, exist_ok=True)
    saved = []
    for name, table in tables.items():
        base = out / f"{prefix}_{name}"
        table.to_csv(f"{base}.csv")
        saved.append(f"{base}.csv")
        with open(f"{base}.tex", "w") as f:
            f.write(table.to_latex(escape=False, caption=f"Table: {name}", label=f"tab:{prefix}_{name}"))
        saved.append(f"{base}.tex")
        try:
            md = table.to_markdown()
        except ImportError:
            flat = table.reset_index()
            cols = [" / ".join(map(str, c)) if isinstance(c, tuple) else str(c) for c in flat.columns]
            lines = ["| " + " | ".join(cols) + " |", "|" + "---|" * len(cols)]
            lines += ["| " + " | ".join(map(str, row)) + " |" for row in flat.itertuples(index=False)]
            md = "\n".join(lines)
        with open(f"{base}.md", "w") as f:
            f.write(md)
        saved.append(f"{base}.md")
    return saved
