#!/usr/bin/env Rscript
# tools/r_parity.R — golden vectors from the REFERENCE ITSELF for the HMM hot path.
#
# R is not available in this project's containers, so the CPU oracle (oracle/hb_oracle.c) is a
# restatement that could only be pinned on what the reference renders (SNP counts, vignette regions).
# This script closes that gap on any machine that has R: it runs the four literal call sites of the
# path on the bundled MGH31 data and dumps, for every (height, group), the exact inputs and outputs
# of `HiddenMarkov::dthmm` -> `HiddenMarkov::Viterbi` (+ `HiddenMarkov::forwardback` posteriors):
#
#   expression   R/HoneyBADGER.R:1122-1151 (method) == R/HoneyBADGER_gexp.R:476-504 (function)
#   allele       R/HoneyBADGER.R:1378-1410 (method) == R/HoneyBADGER_allele.R:553-580 (function)
#
#   Rscript tools/r_parity.R /path/to/HoneyBADGER  tests/golden/r_parity_mgh31.txt
#
# (first argument: a checkout of JEFworks-Lab/HoneyBADGER, only its data/*.rda are read; packages needed:
# HiddenMarkov, caTools.)  tests/test_r_golden.py loads the file when it is present and requires
#   * pooled vectors (apply(.,1,mean), sd, rowSums(.>0)) bit-identical to the oracle's and the device's,
#   * Viterbi paths identical, posteriors within 1e-9,
#   * dnorm / dbinom log-densities of every distinct input bit-identical (the LUT and the literal kernel),
#   * cutree labels identical to the device grouping.
# All doubles are written with 17 significant digits (round-trip exact).

args <- commandArgs(trailingOnly = TRUE)
if (length(args) < 2) stop("usage: Rscript r_parity.R <HoneyBADGER checkout> <out.txt>")
ref <- args[1]
out <- file(args[2], "w")
suppressPackageStartupMessages({ library(HiddenMarkov); library(caTools) })

num <- function(v) paste(sprintf("%.17g", as.numeric(v)), collapse = " ")
int <- function(v) paste(as.integer(v), collapse = " ")
put <- function(...) writeLines(paste0(...), out)

for (f in c("gexp", "ref", "r", "cov.sc")) load(file.path(ref, "data", paste0(f, ".rda")))
put("# hb_r_parity v1  R ", R.version.string, "  HiddenMarkov ", as.character(packageVersion("HiddenMarkov")),
    "  caTools ", as.character(packageVersion("caTools")))

## ------------------------------------------------------------------ expression (3-state Gaussian)
## setGexpMats(filter=FALSE, scale=FALSE) arithmetic (R/HoneyBADGER.R:160-161); gene coordinates come
## from biomaRt (network) in the reference — the stored row order is used instead, which does not
## change any per-call arithmetic.
gexp.norm <- gexp - ref                       # ref is rowMeans of the reference cells, same row order
dev <- 1.015658                               # setGexpDev's value on this data (docs/Getting_Started.md:108)
t <- 1e-6
min.traverse <- 5
k <- 101
mat.smooth <- apply(gexp.norm, 2, caTools::runmean, k)
d <- dist(t(mat.smooth)); d[is.na(d)] <- 0; d[is.nan(d)] <- 0; d[is.infinite(d)] <- 0
hc <- hclust(d, method = "ward.D2")
put("GEXP genes ", nrow(gexp.norm), " cells ", ncol(gexp.norm), " dev ", num(dev), " t ", num(t))
put("RUNMEAN_COL1 ", num(mat.smooth[, 1]))
put("HCLUST_HEIGHT ", num(hc$height))
heights <- seq_len(min(min.traverse, ncol(gexp.norm)))
for (h in heights) {
    ct <- cutree(hc, k = h)
    put("CUTREE gexp h ", h, " labels ", int(ct))
    for (group in unique(ct)) {
        if (sum(ct == group) > 1) {
            x <- apply(gexp.norm[, ct == group], 1, mean)
            s <- sd(x)
            Pi <- matrix(c(1 - 2 * t, t, t, t, 1 - 2 * t, t, t, t, 1 - 2 * t), byrow = TRUE, nrow = 3)
            z <- HiddenMarkov::dthmm(x, Pi, c(0, 1, 0), "norm", list(mean = c(-dev, 0, dev), sd = c(s, s, s)))
            y <- HiddenMarkov::Viterbi(z)
            fb <- HiddenMarkov::forwardback(x, Pi, c(0, 1, 0), "norm", list(mean = c(-dev, 0, dev), sd = c(s, s, s)))
            g <- exp(fb$logalpha + fb$logbeta - fb$LL)
            put("CASE gexp h ", h, " group ", group, " ncell ", sum(ct == group), " n ", length(x))
            put("sd ", num(s))
            put("x ", num(x))
            put("logb1 ", num(dnorm(x, -dev, s, log = TRUE)))
            put("logb2 ", num(dnorm(x, 0, s, log = TRUE)))
            put("logb3 ", num(dnorm(x, dev, s, log = TRUE)))
            put("y ", int(y))
            put("LL ", num(fb$LL))
            for (kk in 1:3) put("gamma", kk, " ", num(g[, kk]))
        }
    }
}

## ------------------------------------------------------------------ allele (2-state binomial)
## setAlleleMats(r, cov.sc, het.deviance.threshold = 0.1, min.cell = 3) arithmetic, R/HoneyBADGER.R:493-595
r.init <- r; n.sc.init <- cov.sc
l <- rowSums(r.init > 0); n.bulk <- rowSums(n.sc.init > 0)
E <- l / n.bulk
vi <- E > 0.1 & E < 0.9
put("ALLELE het ", sum(vi, na.rm = TRUE))
r.init <- r.init[which(vi), ]; n.sc.init <- n.sc.init[which(vi), ]; l <- l[which(vi)]; n.bulk <- n.bulk[which(vi)]
vi <- rowSums(n.sc.init > 0) >= 3
r.init <- r.init[vi, ]; n.sc.init <- n.sc.init[vi, ]; l <- l[vi]; n.bulk <- n.bulk[vi]
E <- l / n.bulk
flip <- E > 0.5
r.maf <- r.init; r.maf[flip, ] <- n.sc.init[flip, ] - r.init[flip, ]
n.sc <- n.sc.init
put("ALLELE snps ", nrow(r.maf), " cells ", ncol(r.maf))
put("SNPNAMES ", paste(rownames(r.maf), collapse = " "))
pd <- 0.1; pn <- 0.45; min.traverse <- 3
mat.tot <- r.maf / n.sc
mat.smooth <- apply(mat.tot, 2, caTools::runmean, k = 31)
d <- dist(t(mat.smooth)); d[is.na(d)] <- 0; d[is.nan(d)] <- 0; d[is.infinite(d)] <- 0
hc <- hclust(d, method = "ward.D2")
put("HCLUST_HEIGHT ", num(hc$height))
heights <- 1:min(min.traverse, ncol(r.maf))
for (h in heights) {
    ct <- cutree(hc, k = h)
    put("CUTREE allele h ", h, " labels ", int(ct))
    for (group in unique(ct)) {
        if (sum(ct == group) > 1) {
            mafl <- rowSums(r.maf[, ct == group] > 0)
            sizel <- rowSums(n.sc[, ct == group] > 0)
            Pi <- matrix(c(1 - t, t, t, 1 - t), byrow = TRUE, nrow = 2)
            z <- HiddenMarkov::dthmm(mafl, Pi, c(0, 1), "binom", list(prob = c(pd, pn)), list(size = sizel), discrete = TRUE)
            y <- HiddenMarkov::Viterbi(z)
            fb <- HiddenMarkov::forwardback(mafl, Pi, c(0, 1), "binom", list(prob = c(pd, pn)), list(size = sizel))
            g <- exp(fb$logalpha + fb$logbeta - fb$LL)
            put("CASE allele h ", h, " group ", group, " ncell ", sum(ct == group), " n ", length(mafl))
            put("mafl ", int(mafl))
            put("sizel ", int(sizel))
            put("logb1 ", num(dbinom(mafl, sizel, pd, log = TRUE)))
            put("logb2 ", num(dbinom(mafl, sizel, pn, log = TRUE)))
            put("y ", int(y))
            put("LL ", num(fb$LL))
            for (kk in 1:2) put("gamma", kk, " ", num(g[, kk]))
        }
    }
}

## ------------------------------------------------------------------ density tables (the emission LUT)
## every (x, n) with n <= 64 plus a sample of deep counts, both states
put("DBINOM_TABLE nmax 64")
for (n in 0:64) put("dbinom n ", n, " pd ", num(dbinom(0:n, n, pd, log = TRUE)), " pn ", num(dbinom(0:n, n, pn, log = TRUE)))
deep <- c(100, 137, 256, 511, 512, 1000, 1245)
for (n in deep) {
    xs <- unique(round(c(0, 1, n * 0.05, n * 0.1, n * 0.3, n * 0.45, n * 0.5, n - 1, n)))
    put("dbinom_deep n ", n, " x ", int(xs), " pd ", num(dbinom(xs, n, pd, log = TRUE)), " pn ", num(dbinom(xs, n, pn, log = TRUE)))
}
close(out)
cat("wrote", args[2], "\n")
