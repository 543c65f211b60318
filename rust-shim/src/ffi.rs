//! Hand-written declarations of include/graph_witness.h (what bindgen generates in the reference, build.rs:9-26).
#![allow(non_camel_case_types)]
use std::ffi::{c_char, c_int, c_void};

pub const GW_ERROR_CODE_OK: c_int = 0;
pub const GW_ERROR_CODE_ERROR: c_int = 1;

#[repr(C)]
pub struct gw_status_t {
    pub code: c_int,
    pub error_msg: *mut c_char,
}

#[repr(C)]
pub struct gw_graph_t {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct gw_graph_info_t {
    pub n_nodes: u64,
    pub n_ops: u64,
    pub n_inputs: u32,
    pub n_witness: u32,
    pub n_input_signals: u32,
    pub n_instrs: u32,
    pub n_regs: u32,
    pub n_spill: u32,
    pub n_mul: u64,
    pub n_div: u64,
    pub n_spill_ld: u64,
    pub n_spill_st: u64,
    pub n_slots: u32,
    pub n_dot: u32,
    pub n_dot_mac: u32,
    pub n_mul_instr: u32,
    pub n_inversions: u32,
    pub threads: u32,
    pub sets_per_thread: u32,
    pub n_narrow_instr: u32,
    pub bit_eligible: u32,
    pub bit_luts: u32,
    pub bit_steps: u32,
    pub bit_wide: u32,
}

pub type gw_witness_chunk_fn = unsafe extern "C" fn(
    user: *mut c_void, device: c_int, first_set: usize, n_sets: usize, rows: *const u8, row_bytes: usize, flags: *const u32,
) -> c_int;

extern "C" {
    pub fn gw_calc_witness(inputs: *const c_char, graph_data: *const c_void, graph_data_len: usize,
                           wtns_data: *mut *mut c_void, wtns_len: *mut usize, status: *const gw_status_t) -> c_int;
    pub fn gw_graph_load(graph_data: *const c_void, graph_data_len: usize, graph: *mut *mut gw_graph_t,
                         status: *mut gw_status_t) -> c_int;
    pub fn gw_graph_free(graph: *mut gw_graph_t);
    pub fn gw_graph_info(graph: *const gw_graph_t, info: *mut gw_graph_info_t) -> c_int;
    pub fn gw_graph_input_signal(graph: *const gw_graph_t, i: u32, name: *mut *const c_char, offset: *mut u32,
                                 len: *mut u32) -> c_int;
    pub fn gw_calc_witness_batch(graph: *mut gw_graph_t, inputs: *const u8, n_sets: usize, witness: *mut u8,
                                 flags: *mut u32, n_gpus: c_int, status: *mut gw_status_t) -> c_int;
    pub fn gw_calc_witness_batch_stream(graph: *mut gw_graph_t, first_device: c_int, n_gpus: c_int, inputs: *const u8,
                                        n_sets: usize, chunk_sets: usize, f: gw_witness_chunk_fn, user: *mut c_void,
                                        status: *mut gw_status_t) -> c_int;
    pub fn gw_wtns_header(n_witness: u32, dst76: *mut u8);
}
