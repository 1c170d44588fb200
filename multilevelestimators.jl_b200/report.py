"""Console output of a verbose run: the tables and messages of src/print.jl, line for line (`verbose = true`).

Purely cosmetic and off by default in this mirror; kept so that a user of the reference sees the familiar progress report.
All functions write to `file` (default sys.stdout)."""
import datetime
import sys

L, N = 81, 15   # l(), n() of src/print.jl:11-15


def short(x):
    return "%5.3f" % x


def shorte(x):
    return "%5.3e" % x


def long(x):
    return "%12.5e" % x


def pprint(index):
    return "(" + ", ".join(str(i) for i in index) + ")"   # src/index.jl:28


class Reporter:
    def __init__(self, est, file=None):
        self.est, self.file = est, file

    # -- primitives (src/print.jl:17-29) -------------------------------------------------------------------------------
    def out(self, s="", end="\n"):
        print(s, end=end, file=self.file or sys.stdout)

    def end_table_row(self, s):
        self.out(s + " " * (L - len(s) - 1) + "|")

    def hline(self):
        self.out("+" + "-" * (L - 2) + "+")

    def table_hline(self, m):
        self.out("+" + ("-" * N + "+") * m)

    def table_header(self, names):
        self.table_hline(len(names))
        self.out("| " + "".join(name + " " * (N - len(name) - 1) + "| " for name in names))
        self.table_hline(len(names))

    def elname(self):   # src/print.jl:88-92
        return "level" if self.est.index_set.ndims == 1 else "index"

    def nb_of_samples_str(self, index, n=None):   # src/print.jl:123-127
        n = self.est.nb_of_samples[index] if n is None else n
        return "%d x %d" % (self.est.options["nb_of_shifts"](index), n) if self.est.qmc else str(n)

    def cell(self, s):
        return " " + s + " " * (N - len(s) - 2) + " |"

    def num(self, x):
        return long(x) + " " * (N - 12 - 1) + " |"

    # -- header / footer (src/print.jl:34-54) --------------------------------------------------------------------------
    def header(self, eps):
        e = self.est
        self.hline()
        self.end_table_row("| *** MultilevelEstimators.jl @" + datetime.datetime.now().isoformat(timespec="milliseconds"))
        self.end_table_row("| *** This is a Estimator{%r, %s}" % (e.index_set, type(e.sample_method).__name__))
        self.end_table_row("| *** Simulating " + e.options["name"].split(".")[0])
        self.end_table_row("| *** Tolerance on RMSE ϵ = " + shorte(eps))
        self.hline()

    def footer(self):
        self.hline()
        self.end_table_row("| *** MultilevelEstimators.jl @" + datetime.datetime.now().isoformat(timespec="milliseconds"))
        self.end_table_row("| *** Successfull termination.")
        self.hline()

    # -- status table (src/print.jl:59-83) -----------------------------------------------------------------------------
    def status(self):
        e = self.est
        self.table_header([self.elname(), "E", "dE", "V", "dV", "N", "W"])
        for index in e.keys():
            if e.nb_of_samples[index] > 0:
                s = "| " + pprint(index) + " " * (N - len(pprint(index)) - 2) + " |"
                s += self.num(e.mean0(index)) + self.num(e.mean(index)) + self.num(e.var0(index)) + self.num(e.var(index))
                s += self.cell(self.nb_of_samples_str(index)) + self.num(e.cost(index))
                self.out(s)
        self.table_hline(7)

    def two_column_table(self, title, rows):
        self.table_header([self.elname(), title])
        for index, text in rows:
            self.out("| " + pprint(index) + " " * (N - len(pprint(index)) - 2) + " |" + self.cell(text))
        self.table_hline(2)

    def optimal_nb_of_samples(self, samples):   # src/print.jl:97-121 (`estimator isa U` there is always false: "to")
        self.out("Samples will be updated to")
        rows = [(i, self.nb_of_samples_str(i, samples[i])) for i in sorted(samples, key=lambda t: t[::-1]) if samples[i] > 0]
        self.two_column_table("N", rows)

    def largest_profit(self, max_index, max_profit, indices, profits):   # src/print.jl:132-151
        self.out("Profit indicators:")
        self.two_column_table("profit", [(i, shorte(p)) for i, p in zip(indices, profits)])
        self.out("Index with max profit is %s (value = %s)." % (pprint(max_index), long(max_profit)))

    def pmf(self):   # src/print.jl:156-175
        self.out("Using approximate probability mass function:")
        P = self.est.pmf
        self.two_column_table("P", [(i, shorte(P[i])) for i in sorted(P, key=lambda t: t[::-1])])

    # -- convergence messages (src/print.jl:180-209) -------------------------------------------------------------------
    def convergence(self, converged):
        self.status()
        if converged:
            self.out("Convergence reached. RMSE ≈" + long(self.est.rmse()) + ".")
        self.footer()

    def rates(self):   # src/print.jl:227-236
        e = self.est

        def one(v):
            return short(v[0]) if e.index_set.ndims == 1 else "(" + ", ".join(short(x) for x in v) + ")"
        self.out("  ==> Rates: α ≈ %s, β ≈ %s, γ ≈ %s." % (one(e.alpha()), one(e.beta()), one(e.gamma())))

    def mse_analysis(self, eps, theta):
        e = self.est
        self.out("Checking convergence...")
        self.rates()
        self.out("  ==> Variance of the estimator ≈" + long(e.varest()) + ".")
        self.out("  ==> Bias of the estimator ≈" + long(e.bias()) + ".")
        self.out("  ==> MSE splitting parameter ≈ " + short(theta) + ".")
        if not e.converged(eps, theta):
            self.out("No convergence yet. RMSE ≈" + long(e.rmse()) + " > " + shorte(eps) + ".")
            self.out("Adding an extra level...")

    def qmc_convergence(self, eps, theta):
        v = self.est.varest()
        conv = not (v > theta * eps**2)
        self.out("Checking if variance of estimator is larger than θ*ϵ^2...")
        self.out("  ==> %s (%s%s%s *%s)" % ("no!" if conv else "yes", long(v)[1:], " < " if conv else " > ", short(theta), long(eps**2)))

    def unbiased_convergence(self, eps):
        v = self.est.varest()
        conv = not (v > eps**2)
        self.out("Checking if variance of estimator is larger than ϵ^2...")
        self.out("  ==> %s (%s%s%s)" % ("no!" if conv else "yes", long(v)[1:], " < " if conv else " > ", long(eps**2)[1:]))
        self.rates()

    # -- sampling, levels, index sets (src/print.jl:214-282) -----------------------------------------------------------
    def sample_header(self, index, n, warm_up):
        self.out("Taking %s%s sample%s at %s %s..." % (self.nb_of_samples_str(index, n), " warm-up" if warm_up else " additional",
                                                      "s" if n > 1 else "", self.elname(), pprint(index)), end="")

    def sample_footer(self):
        self.out(" done")

    def level(self, Lcur):
        from .host import ML, SL
        if isinstance(self.est.index_set, SL):
            self.out("Currently running on finest level.")
        elif isinstance(self.est.index_set, ML):
            self.out("Currently running on level %d." % Lcur)
        else:
            self.out("Currently running with L = %d." % Lcur)

    def index_set(self, index_set):
        from .host import format_index_set
        e = self.est
        if e.index_set.ndims != 2:
            return
        if e.adaptive:
            self.out("Shape of the index set (▣ = active set):")
            self.out(format_index_set(sorted(e.old_set), sorted(e.active_set)), end="")
        elif e.unbiased:
            self.out("Shape of the index set:")
            self.out(format_index_set([i for i in e.keys() if e.has_samples_at_index(i)]), end="")
        else:
            self.out("Shape of the index set:")
            self.out(format_index_set(sorted(set(e.keys()) | set(index_set))), end="")
