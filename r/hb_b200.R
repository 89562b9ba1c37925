## hb_b200.R — the reference-side change: the four `dthmm -> Viterbi` call sites go through .Call.
##
## Drop-in: source this after library(HoneyBADGER) (or patch the package) — signatures, defaults,
## fields and return values of calcGexpCnvBoundaries / calcAlleleCnvBoundaries stay exactly as in
## R/HoneyBADGER.R:1092, :1337, R/HoneyBADGER_gexp.R:462, R/HoneyBADGER_allele.R:543.
## Not executable in this repository's environment (no R); see INTEGRATION.md.

dyn.load("hbb200.so")  # built from r/hb_shim.c against libhb_b200.so

## 1:1 replacements of the HiddenMarkov pair ------------------------------------------------------
hb_viterbi_norm <- function(x, Pi, delta, pm) {
  ## was: HiddenMarkov::Viterbi(HiddenMarkov::dthmm(x, Pi, delta, "norm", pm))
  .Call("C_hb_viterbi_norm", x, as.double(pm$mean), as.double(pm$sd[1]), Pi, as.double(delta))
}
hb_viterbi_binom <- function(x, Pi, delta, pm, pn) {
  ## was: HiddenMarkov::Viterbi(HiddenMarkov::dthmm(x, Pi, delta, "binom", pm, pn, discrete=TRUE))
  .Call("C_hb_viterbi_binom", as.integer(x), as.integer(pn$size), as.double(pm$prob), Pi, as.double(delta))
}

## States + posteriors (gamma[t, chain, state] = P(state_t | all x), HiddenMarkov::forwardback semantics) + LL
hb_posteriors_norm <- function(x, Pi, delta, pm) {
  r <- .Call("C_hb_posteriors_norm", x, as.double(pm$mean), as.double(pm$sd[1]), Pi, as.double(delta))
  list(y = r[[1]], gamma = array(r[[2]], dim = c(NROW(x), NCOL(x), 3L)), LL = r[[3]])
}
hb_posteriors_binom <- function(x, Pi, delta, pm, pn) {
  size <- pn$size
  storage.mode(x) <- "integer"; storage.mode(size) <- "integer"   # keeps a matrix a matrix (as.integer would not)
  r <- .Call("C_hb_posteriors_binom", x, size, as.double(pm$prob), Pi, as.double(delta))
  list(y = r[[1]], gamma = array(r[[2]], dim = c(NROW(x), NCOL(x), 2L)), LL = r[[3]])
}

## R/HoneyBADGER.R:1150-1151 and R/HoneyBADGER_gexp.R:503-504 become
##   results <- hb_viterbi_norm(mat.smooth, Pi, delta, list(mean=c(pd, pn, pa), sd=c(sd,sd,sd)))
## R/HoneyBADGER.R:1409-1410 and R/HoneyBADGER_allele.R:579-580 become
##   results <- hb_viterbi_binom(mafl, Pi, delta, list(prob=c(pd, pn)), list(size=sizel))

## Zero-edit drop-in ------------------------------------------------------------------------------
## The reference reaches the path through `HiddenMarkov::dthmm(...)` followed by
## `HiddenMarkov::Viterbi(z)` at four sites (R/HoneyBADGER.R:1150-1151, :1409-1410,
## R/HoneyBADGER_gexp.R:503-504, R/HoneyBADGER_allele.R:579-580).  `::` resolves the binding at call
## time, so swapping the pair inside the HiddenMarkov namespace routes all four through the GPU without
## touching a line of the package: every signature, default, RefClass field and vignette script stays
## exactly as it is.  Models the library does not cover fall through to the original functions.
.hb_orig <- new.env()
hb_dthmm <- function(x, Pi, delta, distn, pm, pn = NULL, discrete = NULL, nonstat = FALSE) {
  z <- list(x = x, Pi = Pi, delta = delta, distn = distn, pm = pm, pn = pn, discrete = discrete,
            nonstat = nonstat)
  class(z) <- "dthmm"
  z
}
hb_Viterbi <- function(object, ...) {
  on_gpu <- !isTRUE(object$nonstat) && is.matrix(object$Pi) &&
    ((identical(object$distn, "norm") && nrow(object$Pi) == 3L && length(unique(object$pm$sd)) == 1L) ||
     (identical(object$distn, "binom") && nrow(object$Pi) == 2L))
  if (!on_gpu) {
    z <- .hb_orig$dthmm(object$x, object$Pi, object$delta, object$distn, object$pm, object$pn,
                        object$discrete, object$nonstat)
    return(.hb_orig$Viterbi(z, ...))
  }
  ## stop("Problems With Underflow") comes back from the shim with HiddenMarkov's own text, so the
  ## tryCatch around the recursive calls (R/HoneyBADGER.R:1303-1311) behaves as before
  if (identical(object$distn, "norm"))
    hb_viterbi_norm(object$x, object$Pi, object$delta, object$pm)
  else
    hb_viterbi_binom(object$x, object$Pi, object$delta, object$pm, object$pn)
}
hb_b200_install <- function() {
  ns <- asNamespace("HiddenMarkov")
  if (is.null(.hb_orig$dthmm)) {
    .hb_orig$dthmm <- get("dthmm", envir = ns)
    .hb_orig$Viterbi <- get("Viterbi", envir = ns)
  }
  utils::assignInNamespace("dthmm", hb_dthmm, ns = "HiddenMarkov")
  utils::assignInNamespace("Viterbi", hb_Viterbi, ns = "HiddenMarkov")
  invisible(TRUE)
}
hb_b200_uninstall <- function() {
  if (!is.null(.hb_orig$dthmm)) {
    utils::assignInNamespace("dthmm", .hb_orig$dthmm, ns = "HiddenMarkov")
    utils::assignInNamespace("Viterbi", .hb_orig$Viterbi, ns = "HiddenMarkov")
  }
  invisible(TRUE)
}
## usage:  library(HoneyBADGER); source("hb_b200.R"); hb_b200_install()
##         hb$calcGexpCnvBoundaries(init=TRUE); hb$calcAlleleCnvBoundaries(init=TRUE)   # unchanged calls

## Whole-round form (pooling on the GPU too): replaces the `lapply(heights, ...)` block ------------
hb_gexp_round <- function(gexp.norm, hc, heights, dev, t) {
  cts <- lapply(heights, function(h) cutree(hc, k = h))
  member <- do.call(cbind, unlist(lapply(cts, function(ct)
    lapply(unique(ct), function(g) if (sum(ct == g) > 1) ct == g)), recursive = FALSE))
  Pi <- matrix(c(1-2*t, t, t, t, 1-2*t, t, t, t, 1-2*t), byrow = TRUE, nrow = 3)
  y <- .Call("C_hb_gexp_pooled_viterbi", gexp.norm, member, c(-dev, 0, dev), Pi, c(0, 1, 0))
  lapply(seq_len(ncol(y)), function(k) list(amp = rownames(gexp.norm)[y[, k] == 3],
                                            del = rownames(gexp.norm)[y[, k] == 1]))
}

## Grouping and retest on the GPU (optional; SURVEY.md 8f-1/8f-2) ----------------------------------
## R/HoneyBADGER.R:1122-1136: runmean + dist + hclust("ward.D2") + cutree for every height at once
hb_group_gexp <- function(gexp.norm, heights, k = 101L) {
  ct <- .Call("C_hb_group_gexp", gexp.norm, as.integer(k), as.integer(heights))
  rownames(ct) <- colnames(gexp.norm)
  ct   # ct[, i] is cutree(hc, k = heights[i])
}
## R/HoneyBADGER.R:1378-1397
hb_group_allele <- function(r.maf, n.sc, heights, k = 31L) {
  ct <- .Call("C_hb_group_allele", r.maf, n.sc, as.integer(k), as.integer(heights))
  rownames(ct) <- colnames(r.maf)
  ct
}
## deterministic stand-in for calcGexpCnvProb's JAGS run (R/HoneyBADGER.R:374-470): same return shape
hb_calcGexpCnvProb <- function(gexp.norm.sub, m, sigma0 = NULL) {
  p <- .Call("C_hb_retest_gexp", gexp.norm.sub, as.double(m), sigma0)
  for (i in 1:3) names(p[[i]]) <- colnames(gexp.norm.sub)
  names(p) <- c("posterior probability of amplification", "posterior probability of deletion",
                "estimated mean normalized expression deviation")
  p
}

## deterministic stand-in for calcAlleleCnvProb's JAGS run (R/HoneyBADGER.R:907-1042): same 1 x K matrix
hb_calcAlleleCnvProb <- function(r.sub, n.sc.sub, l.sub, n.bulk.sub, geneFactor = NULL, pe = 0.1, mono = 0.7) {
  gf <- if (is.null(geneFactor)) NULL else as.integer(factor(geneFactor[rownames(r.sub)],
                                                            levels = unique(geneFactor[rownames(r.sub)])))
  pm <- .Call("C_hb_retest_allele", r.sub, n.sc.sub, as.double(l.sub), as.double(n.bulk.sub), gf,
              as.double(pe), as.double(mono))
  colnames(pm) <- colnames(r.sub)
  pm
}

## Input construction on the GPU (optional; SURVEY.md 8f-3) ----------------------------------------
## R/HoneyBADGER.R:139-161 inside setGexpMats: everything between the name intersection and the biomaRt
## lookup.  Bit-identical to rowMeans / scale / `-` of base R; returns gexp.norm with dimnames.
hb_setGexpMats_core <- function(gexp.sc, gexp.ref, filter = TRUE, minMeanBoth = 0, minMeanTest = NULL,
                                minMeanRef = NULL, scale = TRUE) {
  gexp.ref <- as.matrix(gexp.ref)
  res <- .Call("C_hb_set_gexp_mats", gexp.sc, gexp.ref, filter, as.double(minMeanBoth), minMeanTest,
               minMeanRef, scale)
  gexp.norm <- res[[1]]
  dimnames(gexp.norm) <- list(rownames(gexp.sc)[res[[2]]], colnames(gexp.sc))
  gexp.norm
}
## R/HoneyBADGER.R:493-595 inside setAlleleMats (the nested lapply of :538-565 is the reference's slowest
## input step): list(r.maf, n.sc, l, n.bulk, l.maf) with names, plus n.het for the cat() of :527.
hb_setAlleleMats_core <- function(r.init, n.sc.init, l.init = NULL, n.bulk.init = NULL, filter = TRUE,
                                  het.deviance.threshold = 0.05, min.cell = 3) {
  res <- .Call("C_hb_set_allele_mats", r.init, n.sc.init, l.init, n.bulk.init, filter,
               as.double(het.deviance.threshold), as.integer(min.cell))
  names(res) <- c("r.maf", "n.sc", "l", "n.bulk", "l.maf", "keep", "n.het")
  nm <- rownames(r.init)[res$keep]
  dimnames(res$r.maf) <- dimnames(res$n.sc) <- list(nm, colnames(r.init))
  names(res$l) <- names(res$n.bulk) <- names(res$l.maf) <- nm
  res
}

## Device-resident recursion (SURVEY.md 8b-3) -------------------------------------------------------
## The reference keeps gexp.norm / r.maf / n.sc in RefClass fields and hands sub-matrices down its recursion
## (`gexp.norm[, g1]`, R/HoneyBADGER.R:1304-1309; `r.maf[, g1]`, :1559-1575).  With a device handle the
## matrices cross PCIe ONCE; the handle lives in a new field `.dev` (an external pointer whose finalizer
## frees the device memory) and every level names its genes / cells by index:
##
##   HoneyBADGER$fields(.dev = "list")                       # list(gexp = <externalptr>, allele = <externalptr>)
##   setGexpMats(...)   : .dev$gexp   <<- hb_dev_gexp(gexp.norm)
##   setAlleleMats(...) : .dev$allele <<- hb_dev_counts(r.maf, n.sc)
##
## calcGexpCnvBoundaries keeps its signature (R/HoneyBADGER.R:1092); where it read `gexp.norm.sub` it now
## carries `cells` (column indices into the resident matrix) and `genes` (row indices, minus bound.genes.old),
## and its body uses the four calls below instead of runmean / dist / hclust / apply(mean) / dthmm / Viterbi.
hb_dev_gexp <- function(gexp.norm) .Call("C_hb_upload_gexp", gexp.norm)
hb_dev_counts <- function(r.maf, n.sc) .Call("C_hb_upload_counts", r.maf, n.sc)
hb_dev_free <- function(dev) invisible(.Call("C_hb_free_handle", dev))
hb_dev_bytes <- function(dev) .Call("C_hb_handle_bytes", dev)

## R/HoneyBADGER.R:1122-1136 (window 101) / :1378-1397 (window 31): ct[, i] == cutree(hc, k = heights[i])
hb_dev_group <- function(dev, genes, cells, heights, k) {
  .Call("C_hb_group_dev", dev, genes, cells, as.integer(k), as.integer(heights))
}
## R/HoneyBADGER.R:1135-1161: all (height, group) chains of one level in one call
hb_dev_gexp_round <- function(dev, genes, cells, ct, dev.mag, t) {
  member <- do.call(cbind, unlist(lapply(seq_len(ncol(ct)), function(i)
    lapply(unique(ct[, i]), function(g) if (sum(ct[, i] == g) > 1) ct[, i] == g)), recursive = FALSE))
  Pi <- matrix(c(1-2*t, t, t, t, 1-2*t, t, t, t, 1-2*t), byrow = TRUE, nrow = 3)
  .Call("C_hb_gexp_round_dev", dev, genes, cells, member, c(-dev.mag, 0, dev.mag), Pi, c(0, 1, 0))
}
## R/HoneyBADGER.R:1395-1417
hb_dev_allele_round <- function(dev, snps, cells, ct, pd, pn, t) {
  member <- do.call(cbind, unlist(lapply(seq_len(ncol(ct)), function(i)
    lapply(unique(ct[, i]), function(g) if (sum(ct[, i] == g) > 1) ct[, i] == g)), recursive = FALSE))
  Pi <- matrix(c(1-t, t, t, 1-t), byrow = TRUE, nrow = 2)
  res <- .Call("C_hb_allele_round_dev", dev, snps, cells, member, c(pd, pn), Pi, c(0, 1))
  names(res) <- c("y", "mafl", "sizel")
  res
}
## R/HoneyBADGER.R:1220 (calcGexpCnvProb on gexp.norm[bound.genes.new, cells]) without the sub-matrix copy
hb_dev_retest_gexp <- function(dev, genes, cells, m, sigma0 = NULL) {
  p <- .Call("C_hb_retest_gexp_dev", dev, genes, cells, as.double(m), sigma0)
  names(p) <- c("posterior probability of amplification", "posterior probability of deletion",
                "estimated mean normalized expression deviation")
  p
}
## R/HoneyBADGER.R:1486 (calcAlleleCnvProb on r.maf[bound.snps.new, cells]) on the resident counts
hb_dev_retest_allele <- function(dev, snps, cells, l.maf, n.bulk, geneFactor = NULL, pe = 0.1, mono = 0.7) {
  gf <- if (is.null(geneFactor)) NULL else as.integer(factor(geneFactor, levels = unique(geneFactor)))
  .Call("C_hb_retest_allele_dev", dev, snps, cells, as.double(l.maf), as.double(n.bulk), gf, as.double(pe),
        as.double(mono))
}
