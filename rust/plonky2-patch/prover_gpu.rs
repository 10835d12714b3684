// Replacement bodies for two steps of plonky2 @ f99ed9c `plonky2/src/plonk/prover.rs` (rows N1a and N1c of SURVEY.md 8f).
// NOT COMPILED HERE (no Rust toolchain in this image): a sketch for the maintainer who applies the [patch] recipe of
// INTEGRATION.md; the same C ABI calls are exercised from Python (intmax_zkp_core_b200/prover.py) and C++
// (host/prover_api.hpp), which is what the parity tests run.
use b200zkp_sys as sys;
use plonky2_field::extension::Extendable;
use plonky2_field::polynomial::{PolynomialCoeffs, PolynomialValues};
use plonky2_field::types::Field;
use plonky2_util::log2_strict;

use crate::fri::oracle::CTX;            // the thread-local b200zkp_ctx of oracle_gpu.rs (made pub(crate) there)
use crate::hash::hash_types::RichField;
use crate::iop::witness::MatrixWitness;
use crate::plonk::circuit_data::{CommonCircuitData, ProverOnlyCircuitData};
use crate::plonk::config::GenericConfig;

/// all_wires_permutation_partial_products + the `pop()` / `concat()` that moves every Z to the front: returns the
/// `zs_partial_products` batch of prove() (Z of every challenge, then the partial products challenge by challenge).
fn zs_partial_products<F: RichField + Extendable<D>, C: GenericConfig<D, F = F>, const D: usize>(
    witness: &MatrixWitness<F>,
    betas: &[F],
    gammas: &[F],
    prover_data: &ProverOnlyCircuitData<F, C, D>,
    common_data: &CommonCircuitData<F, C, D>,
) -> Vec<PolynomialValues<F>> {
    let n = common_data.degree();
    let routed = common_data.config.num_routed_wires;
    // column-major u64 buffers: the routed wires are the first columns of the witness matrix, sigmas[i][j] is transposed
    let wires: Vec<u64> = (0..routed).flat_map(|j| (0..n).map(move |i| witness.get_wire(i, j).to_noncanonical_u64())).collect();
    let sigmas: Vec<u64> = (0..routed).flat_map(|j| (0..n).map(move |i| prover_data.sigmas[i][j].to_noncanonical_u64())).collect();
    let k_is: Vec<u64> = common_data.k_is.iter().map(|k| k.to_canonical_u64()).collect();
    let b: Vec<u64> = betas.iter().map(|x| x.to_canonical_u64()).collect();
    let g: Vec<u64> = gammas.iter().map(|x| x.to_canonical_u64()).collect();
    let cols = betas.len() * (1 + common_data.num_partial_products);
    let mut out = vec![0u64; cols * n];
    let rc = CTX.with(|c| unsafe {
        sys::b200zkp_partial_products_and_zs(*c, wires.as_ptr(), sigmas.as_ptr(), log2_strict(n) as u32, routed as u32,
                                             common_data.quotient_degree_factor as u32, k_is.as_ptr(), b.as_ptr(), g.as_ptr(),
                                             betas.len() as u32, out.as_mut_ptr())
    });
    assert_eq!(rc, 0, "b200zkp_partial_products_and_zs failed");
    out.chunks(n).map(|c| PolynomialValues::new(c.iter().map(|&x| F::from_canonical_u64(x)).collect())).collect()
}

/// tail of compute_quotient_polys + the chunking in prove(): quotient values on the coset 7 <w>, natural order, one
/// vector per challenge -> `all_quotient_poly_chunks` for PolynomialBatch::from_coeffs
fn quotient_poly_chunks<F: RichField>(quotient_values: Vec<Vec<F>>, degree: usize) -> Vec<PolynomialCoeffs<F>> {
    let len = quotient_values[0].len();
    let mut flat: Vec<u64> = quotient_values.iter().flat_map(|v| v.iter().map(|x| x.to_noncanonical_u64())).collect();
    let rc = CTX.with(|c| unsafe {
        sys::b200zkp_coset_intt(*c, flat.as_mut_ptr(), log2_strict(len) as u32, quotient_values.len() as u32,
                                F::coset_shift().to_canonical_u64())
    });
    assert_eq!(rc, 0, "b200zkp_coset_intt failed");
    flat.chunks(degree).map(|c| PolynomialCoeffs::new(c.iter().map(|&x| F::from_canonical_u64(x)).collect())).collect()
}
