//! NOT BUILT IN THIS ENVIRONMENT (no Rust toolchain, plonky2 source not vendored) — see INTEGRATION.md.
//! Sketch of the replacement bodies inside a patched plonky2 @ f99ed9c, `plonky2/src/fri/oracle.rs`:
//! the public signatures of PolynomialBatch::from_values / from_coeffs stay as they are; only the bodies
//! call the C ABI.  `copy-back` mode materialises plonky2's own fields so the rest of prove() is untouched.
use b200zkp_sys as sys;
use plonky2_field::fft::FftRootTable;
use plonky2_field::polynomial::{PolynomialCoeffs, PolynomialValues};
use plonky2_field::types::Field;
use plonky2_util::log2_strict;

use crate::fri::oracle::{PolynomialBatch, SALT_SIZE};
use crate::hash::hash_types::{HashOut, RichField};
use crate::hash::merkle_tree::{MerkleCap, MerkleTree};
use crate::plonk::config::GenericConfig;
use crate::util::timing::TimingTree;
use plonky2_field::extension::Extendable;

/// four canonical words of the library's digest layout -> plonky2's HashOut (PoseidonGoldilocksConfig: Hasher::Hash = HashOut<F>)
fn hash_out_from_u64s<F: RichField>(w: &[u64]) -> HashOut<F> {
    HashOut { elements: [F::from_canonical_u64(w[0]), F::from_canonical_u64(w[1]), F::from_canonical_u64(w[2]), F::from_canonical_u64(w[3])] }
}

thread_local! { static CTX: *mut sys::b200zkp_ctx = unsafe {
    let mut c = std::ptr::null_mut();
    assert_eq!(sys::b200zkp_ctx_create(0, std::ptr::null_mut(), &mut c), 0, "no CUDA device (there is no CPU fallback)");
    c
}; }

impl<F: RichField + Extendable<D>, C: GenericConfig<D, F = F>, const D: usize> PolynomialBatch<F, C, D> {
    pub fn from_values(values: Vec<PolynomialValues<F>>, rate_bits: usize, blinding: bool, cap_height: usize,
                       _timing: &mut TimingTree, _fft_root_table: Option<&FftRootTable<F>>) -> Self {
        // GoldilocksField is #[repr(transparent)] over u64: flatten the k columns into one column-major buffer
        let n = values[0].len();
        let flat: Vec<u64> = values.iter().flat_map(|v| v.values.iter().map(|x| x.to_noncanonical_u64())).collect();
        Self::commit_gpu(&flat, false, n, values.len(), rate_bits, blinding, cap_height)
    }

    pub fn from_coeffs(polynomials: Vec<PolynomialCoeffs<F>>, rate_bits: usize, blinding: bool, cap_height: usize,
                       _timing: &mut TimingTree, _fft_root_table: Option<&FftRootTable<F>>) -> Self {
        let n = polynomials[0].len();
        let flat: Vec<u64> = polynomials.iter().flat_map(|p| p.coeffs.iter().map(|x| x.to_noncanonical_u64())).collect();
        Self::commit_gpu(&flat, true, n, polynomials.len(), rate_bits, blinding, cap_height)
    }

    fn commit_gpu(flat: &[u64], is_coeffs: bool, n: usize, k: usize, rate_bits: usize, blinding: bool, cap_height: usize) -> Self {
        let n_log = log2_strict(n);
        let big_n = n << rate_bits;
        let salt: Option<Vec<u64>> = blinding.then(|| (0..SALT_SIZE * big_n).map(|_| F::rand().to_canonical_u64()).collect());
        let mut h = std::ptr::null_mut();
        let rc = CTX.with(|c| unsafe {
            let f = if is_coeffs { sys::b200zkp_commit_from_coeffs } else { sys::b200zkp_commit_from_values };
            f(*c, flat.as_ptr(), n_log as u32, k as u32, rate_bits as u32, cap_height as u32,
              salt.as_ref().map_or(std::ptr::null(), |s| s.as_ptr()), &mut h)
        });
        assert_eq!(rc, 0, "b200zkp commit failed");   // plonky2 panics on the same conditions (asserts)
        // copy-back (strict drop-in): polynomials, leaves, digests, cap -> the existing struct fields
        let row = k + if blinding { SALT_SIZE } else { 0 };
        let mut coeffs = vec![0u64; k * n];
        let mut leaves = vec![0u64; big_n * row];
        let mut digests = vec![0u64; 4 * 2 * (big_n - (1 << cap_height))];
        let mut cap = vec![0u64; 4 << cap_height];
        unsafe {
            sys::b200zkp_batch_coeffs(h, coeffs.as_mut_ptr());
            sys::b200zkp_batch_leaves(h, leaves.as_mut_ptr());
            sys::b200zkp_batch_digests(h, digests.as_mut_ptr());
            sys::b200zkp_batch_cap(h, cap.as_mut_ptr());
            sys::b200zkp_batch_free(h);
        }
        Self {
            polynomials: coeffs.chunks(n).map(|c| PolynomialCoeffs::new(c.iter().map(|&x| F::from_canonical_u64(x)).collect())).collect(),
            merkle_tree: MerkleTree {
                leaves: leaves.chunks(row).map(|r| r.iter().map(|&x| F::from_canonical_u64(x)).collect()).collect(),
                digests: digests.chunks(4).map(hash_out_from_u64s).collect(),
                cap: MerkleCap(cap.chunks(4).map(hash_out_from_u64s).collect()),
            },
            degree_log: n_log, rate_bits, blinding,
        }
    }
}
