"""Result table of the evaluation CLI: same rows, columns and final summary text as the reference's
curses table (``eval/utils/print_table.py:7-90``).  Without a TTY (batch jobs, tests) the live
updates are skipped and only the summary is printed."""
from __future__ import annotations

import sys

import numpy as np


class PrintTable:
    def __init__(self, test_seq_len, row_names, live=None):
        self.test_seq_len = list(test_seq_len)
        self.test_seq_len_sec = ['(' + str(i) + 's)' for i in (np.asarray(test_seq_len) + 1) // 2]
        self.row_names = list(row_names)
        self.line_int = '{:^6}\t' * len(self.test_seq_len)
        self.line_float = '{:>4.2f}\t' * len(self.test_seq_len)
        self.live = sys.stdout.isatty() if live is None else live
        self.rows_cache = None
        self.avg_search_time_cache = float('nan')

    def update_table(self, rows):
        self.rows_cache = rows

    def update_counter(self, i, niter, t):
        self.avg_search_time_cache = t
        if self.live:
            top1 = '' if self.rows_cache is None else ' top1 ' + self.line_float.format(*self.rows_cache[0]).strip()
            print(f'\r{i}/{niter}  {t:>4.2f} ms/query{top1}', end='', flush=True)

    def summary_lines(self):
        cyan, dflt = '\033[36m', '\033[0m'
        lines = ['========= Top1 hit rate (%) of segment-level search =========',
                 ' ' * 14 + ' ' + '{:^43}\t'.format('---------------- Query length ----------------'),
                 '{:^14}'.format('segments') + ' ' + cyan + self.line_int.format(*self.test_seq_len) + ' ' + dflt,
                 '{:^14}'.format('seconds') + ' ' + cyan + self.line_int.format(*self.test_seq_len_sec) + ' ' + dflt,
                 '']
        for i, line in enumerate(self.rows_cache):
            lines.append('{:^14}'.format(self.row_names[i]) + ' ' + self.line_float.format(*line))
        lines.append('=============================================================')
        lines.append(f'average search + evaluation time {self.avg_search_time_cache:>4.2f} ms/query')
        return lines

    def close_table(self):
        if self.live:
            print()
        for ln in self.summary_lines():
            print(ln)
