//! SOURCE ONLY -- compile-unverified (no Rust toolchain in the build image).
//! Raw bindings of include/zkdcap_b200.h; see INTEGRATION.md for the halo2_proofs patch that uses them.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] #[derive(Clone, Copy)] pub struct zk_fr(pub [u64; 4]);
#[repr(C)] #[derive(Clone, Copy)] pub struct zk_g1_affine { pub x: [u64; 4], pub y: [u64; 4] }
#[repr(C)] #[derive(Clone, Copy)] pub struct zk_g1 { pub x: [u64; 4], pub y: [u64; 4], pub z: [u64; 4] }
#[repr(C)] pub struct zk_ctx { _p: [u8; 0] }
#[repr(C)] pub struct zk_domain { _p: [u8; 0] }
#[repr(C)] pub struct zk_pk { _p: [u8; 0] }
#[repr(C)] #[derive(Clone, Copy)] pub struct zk_value_source { pub kind: u32, pub a: u32, pub b: u32 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct zk_calculation { pub kind: u32, pub a: zk_value_source, pub b: zk_value_source, pub parts_off: u32, pub parts_len: u32, pub target: u32 }
#[repr(C)]
pub struct zk_graph {
    pub constants: *const zk_fr, pub n_constants: u32,
    pub rotations: *const i32, pub n_rotations: u32,
    pub calcs: *const zk_calculation, pub n_calcs: u32,
    pub horner_parts: *const zk_value_source, pub n_horner_parts: u32,
    pub num_intermediates: u32,
}
#[repr(C)] #[derive(Clone, Copy)] pub struct zk_column { pub column_type: u32, pub index: u32 }
#[repr(C)]
pub struct zk_evaluator_desc {
    pub k: u32, pub degree: u32, pub blinding_factors: u32,
    pub n_fixed: u32, pub n_advice: u32, pub n_instance: u32, pub n_challenges: u32,
    pub custom_gates: zk_graph, pub lookups: *const zk_graph, pub n_lookups: u32,
    pub perm_columns: *const zk_column, pub n_perm_columns: u32,
}
#[repr(C)] #[derive(Clone, Copy)] pub struct zk_query { pub column: u32, pub rotation: i32 }
#[repr(C)]
pub struct zk_circuit_desc {
    pub ev: zk_evaluator_desc,
    pub lookup_inputs: *const zk_graph, pub lookup_tables: *const zk_graph,
    pub advice_queries: *const zk_query, pub n_advice_queries: u32,
    pub fixed_queries: *const zk_query, pub n_fixed_queries: u32,
    pub transcript_repr: zk_fr,
    pub instance_queries: *const zk_query, pub n_instance_queries: u32,
}
#[repr(C)]
pub struct zk_verifier_params {
    pub fixed_commitments: *const zk_g1_affine, pub permutation_commitments: *const zk_g1_affine,
    pub g1: zk_g1_affine, pub g2: *const u8, pub s_g2: *const u8, pub transcript_kind: c_int,
}
#[repr(C)] pub struct zk_pkey { _p: [u8; 0] }
#[repr(C)] pub struct zk_group { _p: [u8; 0] }
pub const ZK_BASES_HOST: c_int = 0;
pub const ZK_BASES_G: c_int = 1;
pub const ZK_BASES_G_LAGRANGE: c_int = 2;

extern "C" {
    pub fn zk_ctx_create(device: c_int, cuda_stream: *mut c_void, out: *mut *mut zk_ctx) -> c_int;
    pub fn zk_ctx_destroy(ctx: *mut zk_ctx);
    pub fn zk_last_error(ctx: *const zk_ctx) -> *const c_char;
    pub fn zk_srs_upload(ctx: *mut zk_ctx, k: u32, g: *const zk_g1_affine, g_lagrange: *const zk_g1_affine, precompute_windows: c_int) -> c_int;
    pub fn zk_msm(ctx: *mut zk_ctx, scalars: *const zk_fr, bases: *const zk_g1_affine, which: c_int, n: usize, out: *mut zk_g1) -> c_int;
    pub fn zk_ntt(ctx: *mut zk_ctx, a: *mut zk_fr, omega: *const zk_fr, log_n: u32) -> c_int;
    pub fn zk_domain_create(ctx: *mut zk_ctx, j: u32, k: u32, out: *mut *mut zk_domain) -> c_int;
    pub fn zk_domain_destroy(d: *mut zk_domain);
    pub fn zk_lagrange_to_coeff(d: *mut zk_domain, a: *mut zk_fr) -> c_int;
    pub fn zk_coeff_to_lagrange(d: *mut zk_domain, a: *mut zk_fr) -> c_int;
    pub fn zk_coeff_to_extended(d: *mut zk_domain, a: *const zk_fr, out_ext: *mut zk_fr) -> c_int;
    pub fn zk_extended_to_coeff(d: *mut zk_domain, a_ext: *const zk_fr, out: *mut zk_fr) -> c_int;
    pub fn zk_divide_by_vanishing_poly(d: *mut zk_domain, a_ext: *mut zk_fr) -> c_int;
    pub fn zk_pk_upload(ctx: *mut zk_ctx, desc: *const zk_evaluator_desc, fixed: *const *const zk_fr, perm: *const *const zk_fr,
                        l0: *const zk_fr, l_last: *const zk_fr, l_active_row: *const zk_fr, coeff_form: c_int, out: *mut *mut zk_pk) -> c_int;
    pub fn zk_pk_destroy(pk: *mut zk_pk);
    pub fn zk_evaluate_h(pk: *mut zk_pk, advice: *const *const zk_fr, instance: *const *const zk_fr, challenges: *const zk_fr,
                         y: *const zk_fr, beta: *const zk_fr, gamma: *const zk_fr, theta: *const zk_fr,
                         lookups: *const *const zk_fr, perm_products: *const *const zk_fr, out_ext: *mut zk_fr) -> c_int;
    // keygen_vk + keygen_pk / create_proof with every n-sized loop on the device
    pub fn zk_keygen(ctx: *mut zk_ctx, circuit: *const zk_circuit_desc, fixed_values: *const *const zk_fr, perm_mapping: *const u32,
                     out: *mut *mut zk_pkey) -> c_int;
    pub fn zk_pkey_destroy(pk: *mut zk_pkey);
    pub fn zk_pkey_vk_commitments(pk: *const zk_pkey, fixed_commitments: *mut zk_g1_affine, perm_commitments: *mut zk_g1_affine) -> c_int;
    pub fn zk_pkey_pk(pk: *mut zk_pkey) -> *mut zk_pk;
    pub fn zk_pkey_set_transcript(pk: *mut zk_pkey, kind: c_int) -> c_int;      // 0 Blake2b (halo2), 1 EVM / Keccak, 2 Poseidon (snark-verifier)
    pub fn zk_proof_randoms(pk: *const zk_pkey) -> usize;
    pub fn zk_proof_max_bytes(pk: *const zk_pkey) -> usize;
    pub fn zk_create_proof(pk: *mut zk_pkey, advice_values: *const *const zk_fr, instance_values: *const *const zk_fr,
                           instance_lens: *const u32, randoms: *const zk_fr, n_randoms: usize, proof: *mut u8, proof_cap: usize,
                           proof_len: *mut usize) -> c_int;
    // ParamsKZG::{read, write}, G1Affine::{to_bytes, from_bytes} on the device
    pub fn zk_params_read(ctx: *mut zk_ctx, file: *const u8, len: usize, precompute_windows: c_int, k_out: *mut u32,
                          g2: *mut u8, s_g2: *mut u8) -> c_int;
    pub fn zk_params_write(ctx: *mut zk_ctx, g2: *const u8, s_g2: *const u8, file: *mut u8, cap: usize, len: *mut usize) -> c_int;
    pub fn zk_g1_compress(ctx: *mut zk_ctx, points: *const zk_g1_affine, n: usize, out: *mut u8) -> c_int;
    pub fn zk_g1_decompress(ctx: *mut zk_ctx, input: *const u8, n: usize, out: *mut zk_g1_affine) -> c_int;
    /// format: 0 = halo2curves 0.3.1 compressed, 1 = halo2curves >= 0.4 compressed, 2 = SerdeFormat::RawBytes
    pub fn zk_g1_encode(ctx: *mut zk_ctx, format: c_int, points: *const zk_g1_affine, n: usize, out: *mut u8) -> c_int;
    pub fn zk_g1_decode(ctx: *mut zk_ctx, format: c_int, input: *const u8, n: usize, out: *mut zk_g1_affine) -> c_int;
    pub fn zk_params_read_format(ctx: *mut zk_ctx, format: c_int, file: *const u8, len: usize, precompute_windows: c_int,
                                 k_out: *mut u32, g2: *mut u8, s_g2: *mut u8) -> c_int;
    pub fn zk_params_write_format(ctx: *mut zk_ctx, format: c_int, g2: *const u8, s_g2: *const u8, file: *mut u8, cap: usize,
                                  len: *mut usize) -> c_int;
    // one proof over several GPUs (one process per GPU)
    pub fn zk_group_unique_id(id: *mut u8) -> c_int;                 // 256 bytes
    pub fn zk_group_create(ctx: *mut zk_ctx, rank: u32, world: u32, id: *const u8, out: *mut *mut zk_group) -> c_int;
    pub fn zk_group_destroy(g: *mut zk_group);
    pub fn zk_pkey_set_group(pk: *mut zk_pkey, g: *mut zk_group) -> c_int;
    pub fn zk_group_p2p_enabled(g: *const zk_group) -> c_int;
    /// ONE MSM sharded by base range: this rank's slice of the scalars against its resident slice of the bases (device memory)
    pub fn zk_msm_sharded_dev(g: *mut zk_group, d_scalars: *const zk_fr, which: c_int, n_local: usize, d_out: *mut zk_g1) -> c_int;
    pub fn zk_ctx_share_srs(dst: *mut zk_ctx, src: *const zk_ctx) -> c_int;
    pub fn zk_srs_window(ctx: *mut zk_ctx, which: c_int, c_out: *mut u32, w_out: *mut u32) -> c_int;
    // evaluate_h compiled from the expression graph (NVRTC at zk_pk_upload / zk_keygen)
    pub fn zk_pk_evaluator_kind(pk: *const zk_pk) -> c_int;          // 1 = generated kernel, 0 = interpreter
    pub fn zk_evalh_generate_source(desc: *const zk_evaluator_desc, host_flavour: c_int, buf: *mut c_char, cap: usize, len: *mut usize) -> c_int;
    // ParamsKZG.g2 / s_g2: file encoding <-> the 128 uncompressed bytes zk_verify_proof takes
    pub fn zk_poseidon_permute(state: *mut zk_fr, t: u32, r_f: u32, r_p: u32) -> c_int;
    pub fn zk_poseidon_hash(elements: *const zk_fr, count: usize, out: *mut zk_fr, squeezes: u32) -> c_int;
    pub fn zk_g2_decode(format: c_int, input: *const u8, out_uncompressed: *mut u8) -> c_int;
    pub fn zk_g2_encode(format: c_int, in_uncompressed: *const u8, out: *mut u8) -> c_int;
    // verify_proof (host code with its own BN254 pairing)
    pub fn zk_verify_proof(circuit: *const zk_circuit_desc, vp: *const zk_verifier_params, instance_values: *const *const zk_fr,
                           instance_lens: *const u32, proof: *const u8, proof_len: usize, accepted: *mut c_int) -> c_int;
}
