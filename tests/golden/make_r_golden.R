# Writes tests/golden/r_semantics.json from a REAL R session -- run by a maintainer who has R with
# Eutropia (and its dependencies sf, data.table) installed; it cannot run in the build image (no R).
# Same inputs and JSON layout as make_r_semantics.py.  The package's own functions are called where they
# can be called in isolation (get_cellFieldEnv_ex, scavenge_compounds); the statements that live inline in
# run_simulation() / agentFBA_ex() (claim rescale, uptake / production triplets, aggregation and clamp) are
# copied here as statements, with their line numbers, because they cannot be reached without an LP solve.
#
#   Rscript tests/golden/make_r_golden.R          # from the repository root
#
# NOT executed in this repository's CI: if it fails on your R, fix the script, not the fixture.
suppressPackageStartupMessages({ library(data.table); library(sf) })
get_cellFieldEnv_ex <- get("get_cellFieldEnv_ex", asNamespace("Eutropia"))
scavenge_compounds  <- get("scavenge_compounds",  asNamespace("Eutropia"))

num  <- function(x) ifelse(is.na(x), "null", sprintf("%.17g", x))
arr  <- function(x) paste0("[", paste(num(x), collapse = ", "), "]")
iarr <- function(x) paste0("[", paste(as.integer(x), collapse = ", "), "]")
mat  <- function(m) paste0("[", paste(apply(m, 1, arr), collapse = ", "), "]")

# ---- inputs (identical to make_r_semantics.py) --------------------------------------------------------
layer1 <- expand.grid(x = c(3, 1, 0, 2), y = c(2, 0, 1)); layer1$z <- 0          # x fastest, like the Python list
layer2 <- expand.grid(y = c(0.5, 1.5, -0.5), x = c(1.5, -0.5, 2.5, 0.5)); layer2$z <- 1
pts <- rbind(layer1[, c("x", "y", "z")], layer2[, c("x", "y", "z")])
N <- nrow(pts)
fs <- 1
fieldVol <- 16/9 * (fs * 3 / (2 * sqrt(6)))^3 * sqrt(3)                            # R/growthEnvironment.R:101-102
cells <- data.table(x = c(1, 2), y = c(1, 1), size = c(1, 1), scvr = c(1, 1))

gridDT <- as.data.table(pts)                                                       # R/run_simulation.R:165-168
gridDT[, field := 1:.N]
setkeyv(gridDT, c("z", "x", "y"))

pre <- lapply(1:nrow(cells), function(i) {                                          # :217-232
  res <- list(x = cells$x[i], y = cells$y[i], size = cells$size[i], scvr = cells$scvr[i])
  res$gridLC <- gridDT[abs(x - res$x) <= (res$size / 2 + res$scvr) &
                       abs(y - res$y) <= (res$size / 2 + res$scvr) &
                       abs(z - 0)     <= (res$size / 2 + res$scvr)]
  res
})
cellFieldEnv <- lapply(pre, get_cellFieldEnv_ex)                                    # :234, :591-624
acc_raw <- unlist(lapply(cellFieldEnv, function(e) e$acc.prop))

cellFieldEnv <- rbindlist(cellFieldEnv, idcol = TRUE)                               # :236-241
cellFieldEnv[, field_claim_sum := sum(acc.prop), by = "field.id"]
cellFieldEnv[field_claim_sum > 1, acc.prop.tmp := acc.prop^exp(1)]
cellFieldEnv[field_claim_sum > 1, acc.prop.tmp := acc.prop.tmp / sum(acc.prop.tmp) * field_claim_sum, by = "field.id"]
cellFieldEnv[field_claim_sum > 1, acc.prop := acc.prop.tmp / field_claim_sum]
claimed <- sort(unique(cellFieldEnv[field_claim_sum > 1, field.id]))

conc <- matrix(0, N, 4)
i0 <- 0:(N - 1)
conc[, 1] <- 0.5 + 0.25 * (i0 %% 5); conc[, 2] <- 0; conc[, 3] <- 7; conc[, 4] <- 2^-30 * (1 + (i0 %% 3))
untouched <- which(pts$x == -0.5 & pts$y == -0.5 & pts$z == 1)
conc[untouched, 1] <- 5e-13
isConstant <- c(FALSE, FALSE, TRUE, FALSE)
cpds <- paste0("cpd", 1:4)

# ---- gather (R/scavenge_compounds.R:12-31 through the package's own function) --------------------------
acc <- lapply(1:nrow(cells), function(i) {
  le <- cellFieldEnv[.id == i]
  scavenge_compounds(le, conc[le$field.id, , drop = FALSE], cpds, fieldVol)
})
fmol <- do.call(rbind, lapply(acc, function(a) a$fmol))

# ---- adjust_uptake arithmetic (R/adjust_uptake.R:27-32) -------------------------------------------------
deltaTime <- 1/6; cMass <- c(0.5, 0.75); lowbnd <- c(-10, -1000, -1e-9, 0)
ex_lb <- t(sapply(1:2, function(i) -apply(cbind(fmol[i, ] * (1 / deltaTime) / cMass[i], -lowbnd), 1, min)))

# ---- scatter: R/run_simulation.R:682-740 per cell, :309-319 aggregate / apply / clamp -------------------
exchanges <- list(c(cpd1 = -3.0e-4, cpd2 = 2.0e-3, cpd3 = 1.0e-3, cpd4 = -1.0),
                  c(cpd1 = -5.0e-4, cpd2 = -1.0e-3, cpd4 = 4.0e-6))
res <- lapply(1:2, function(i) {
  accCompounds <- acc[[i]]; ex.flx <- exchanges[[i]]
  accmat <- as.matrix(accCompounds$fmolPerField); colnames(accmat) <- accCompounds$compounds
  accmat.row2field <- accCompounds$field.id
  accmat <- t(t(accmat) / colSums(accmat))                                          # :696
  ex.upt <- ex.flx[ex.flx < 0]
  accmat.upt <- accmat[, names(ex.upt), drop = FALSE]
  accmat.upt <- t(t(accmat.upt) * ex.upt)                                           # :707
  accmat.upt <- (accmat.upt / 1e12) / (fieldVol / 1e15)                             # :708
  I.up <- data.table(field.id = rep(accmat.row2field, ncol(accmat.upt)),
                     concMatInd = rep(match(colnames(accmat.upt), cpds), each = nrow(accmat.upt)),
                     concChange = as.numeric(accmat.upt))
  ex.pro <- ex.flx[ex.flx > 0]
  field.frac <- max(accCompounds$field.dist) - accCompounds$field.dist              # :729-730
  field.frac <- field.frac / sum(field.frac)
  ex.pro <- (ex.pro / 1e12) / (fieldVol / 1e15)                                     # :732
  pro.mat <- field.frac %*% t(ex.pro)
  I.pd <- data.table(field.id = rep(accmat.row2field, ncol(pro.mat)),
                     concMatInd = rep(match(colnames(pro.mat), cpds), each = nrow(pro.mat)),
                     concChange = as.numeric(pro.mat))
  list(I.up = I.up, I.pd = I.pd)
})
I <- rbindlist(lapply(res, function(x) rbindlist(list(x$I.up, x$I.pd))))            # :309-319
I <- I[, list(concChange = sum(concChange)), by = c("field.id", "concMatInd")]
I <- I[concMatInd %in% which(!isConstant)]
I <- as.matrix(I)
I[, 3] <- ifelse(is.na(I[, 3]), 0, I[, 3])
after <- conc
after[I[, 1:2]] <- after[I[, 1:2]] + I[, 3]
after[I[, 1:2]] <- ifelse(after[I[, 1:2]] < 1e-12, 0, after[I[, 1:2]])

# ---- correct_numbers (R/growthEnvironment.R:199-217), one axis -------------------------------------------
vals  <- c(0, 0.5, 1, 1.5000000000000002, 2)
all_x <- data.table(s = vals, val = vals); setattr(all_x, "sorted", "s"); setkey(all_x, s)   # as :175, :178
qn <- c(0.5, 0.49999999999999994, 0.96, 0.89, 1.5, 1.5000000000000004, 2.0000000000000004, -0.05, -0.2)
new_n <- all_x[J(qn), roll = -Inf]$val
new_n <- ifelse(abs(new_n - qn) > fs / 10, NA, new_n)

json <- paste0("{\n",
  ' "about": "Known answers for the R-language steps, written by R itself (tests/golden/make_r_golden.R)",\n',
  ' "generated_by": "', R.version.string, ', Eutropia ', as.character(packageVersion("Eutropia")), ', sf ', as.character(packageVersion("sf")),
  ', data.table ', as.character(packageVersion("data.table")), '",\n',
  ' "assumptions": [],\n',
  ' "field_pts": ', mat(as.matrix(pts)), ', "field_vol": ', num(fieldVol), ',\n',
  ' "cells": ', mat(as.matrix(cells)), ',\n',
  ' "cell_ptr": ', iarr(c(0, cumsum(table(factor(cellFieldEnv$.id, levels = 1:2))))), ',\n',
  ' "field_id": ', iarr(cellFieldEnv$field.id), ', "field_dist": ', arr(cellFieldEnv$field.dist), ',\n',
  ' "acc_raw": ', arr(acc_raw), ', "acc_prop": ', arr(cellFieldEnv$acc.prop), ',\n',
  ' "claimed_fields": ', iarr(claimed), ',\n',
  ' "conc": ', mat(conc), ', "is_constant": [', paste(tolower(as.character(isConstant)), collapse = ", "), '],\n',
  ' "fmol": ', mat(fmol), ',\n',
  ' "adjust_uptake": {"delta_time": ', num(deltaTime), ', "c_mass": ', arr(cMass), ', "lowbnd": ', arr(lowbnd), ', "ex_react_lb": ', mat(ex_lb), '},\n',
  ' "exchanges": [', paste(sapply(exchanges, function(e) paste0("[", paste(sprintf("[%d, %s]", match(names(e), cpds), num(e)), collapse = ", "), "]")), collapse = ", "), '],\n',
  ' "after": ', mat(after), ', "touched": ', mat(I[, 1:2, drop = FALSE]), ', "untouched_small_entry": [', untouched, ', 1],\n',
  ' "snap": {"all": ', arr(all_x$val), ', "queries": ', arr(qn), ', "snapped": ', arr(new_n), ', "field_size": 1.0}\n}\n')
writeLines(json, "tests/golden/r_semantics.json")
cat("tests/golden/r_semantics.json written by", R.version.string, "\n")
