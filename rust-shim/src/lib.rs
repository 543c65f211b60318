//! Drop-in for the hot path of iden3/circom-witnesscalc (`calc_witness -> graph::evaluate`) backed by the CUDA library.
//!
//! Same names and signatures as the reference crate (src/lib.rs:19, :114-123, :125-136, :183-247; src/graph.rs:367):
//! `calc_witness`, `wtns_from_witness`, `deserialize_inputs`, `InputSignalsInfo`, `Error`, `graph::evaluate`, plus the
//! batch siblings the reference does not have (`DeviceGraph`, `evaluate_batch`, `evaluate_stream`).
//!
//! UNVERIFIED: written without a Rust toolchain (none exists in the build image); see INTEGRATION.md.
pub mod ffi;
pub mod graph;

use std::collections::HashMap;
use std::ffi::{c_void, CStr, CString};

use ruint::aliases::U256;
use ruint::ParseError;

pub type InputSignalsInfo = HashMap<String, (usize, usize)>;

#[derive(Debug)]
pub enum Error {
    InputsUnmarshal(String),
    InputFieldNumberParseError(ParseError),
    /// anything the device library reports (malformed graph, unknown input name, no GPU, ...): the reference panics
    /// in most of these cases (lib.rs:130,158-161,196)
    Device(String),
}

impl From<ParseError> for Error {
    fn from(e: ParseError) -> Self {
        Error::InputFieldNumberParseError(e)
    }
}

fn take_status(st: &mut ffi::gw_status_t) -> String {
    if st.error_msg.is_null() {
        return String::from("unknown error");
    }
    let s = unsafe { CStr::from_ptr(st.error_msg) }.to_string_lossy().into_owned();
    unsafe { libc::free(st.error_msg as *mut c_void) };
    st.error_msg = std::ptr::null_mut();
    s
}

/// calc_witness (lib.rs:125-136): inputs JSON + graph file bytes -> the witness.  One call = gw_calc_witness, whose
/// .wtns image is unpacked again (76-byte header, then 32-byte little-endian canonical values).
pub fn calc_witness(inputs: &str, graph_data: &[u8]) -> Result<Vec<U256>, Error> {
    let c_inputs = CString::new(inputs).map_err(|e| Error::InputsUnmarshal(e.to_string()))?;
    let mut wtns: *mut c_void = std::ptr::null_mut();
    let mut wtns_len: usize = 0;
    let mut st = ffi::gw_status_t { code: ffi::GW_ERROR_CODE_OK, error_msg: std::ptr::null_mut() };
    let rc = unsafe {
        ffi::gw_calc_witness(c_inputs.as_ptr(), graph_data.as_ptr() as *const c_void, graph_data.len(), &mut wtns, &mut wtns_len, &st)
    };
    if rc != 0 {
        return Err(Error::Device(take_status(&mut st)));
    }
    let bytes = unsafe { std::slice::from_raw_parts(wtns as *const u8, wtns_len) };
    let witness = bytes[76..].chunks_exact(32).map(U256::from_le_slice).collect();
    unsafe { libc::free(wtns) };
    Ok(witness)
}

/// wtns_from_witness (lib.rs:114-123): snarkjs .wtns v2 framing; the header comes from the library (gw_wtns_header)
pub fn wtns_from_witness(witness: Vec<U256>) -> Vec<u8> {
    let mut buf = vec![0u8; 76 + 32 * witness.len()];
    unsafe { ffi::gw_wtns_header(witness.len() as u32, buf.as_mut_ptr()) };
    for (i, v) in witness.iter().enumerate() {
        buf[76 + 32 * i..108 + 32 * i].copy_from_slice(&v.to_le_bytes::<32>());
    }
    buf
}

/// deserialize_inputs (lib.rs:195-247): same accepted forms (decimal string, non-negative integer, flat array of both)
pub fn deserialize_inputs(inputs_data: &[u8]) -> Result<HashMap<String, Vec<U256>>, Error> {
    let v: serde_json::Value = serde_json::from_slice(inputs_data).map_err(|e| Error::InputsUnmarshal(e.to_string()))?;
    let map = match v {
        serde_json::Value::Object(m) => m,
        _ => return Err(Error::InputsUnmarshal("inputs must be an object".to_string())),
    };
    fn scalar(k: &str, v: &serde_json::Value, in_array: bool) -> Result<U256, Error> {
        match v {
            serde_json::Value::String(s) => Ok(U256::from_str_radix(s, 10)?),
            serde_json::Value::Number(n) => n
                .as_u64()
                .map(U256::from)
                .ok_or_else(|| Error::InputsUnmarshal(format!("signal value is not a positive integer: {}", k))),
            _ if in_array => Err(Error::InputsUnmarshal(format!("inputs must be a string: {}", k))),
            _ => Err(Error::InputsUnmarshal(format!("value for key {} must be an a number as a string, as a number of an array of strings of numbers", k))),
        }
    }
    let mut inputs = HashMap::new();
    for (k, v) in map {
        let vals = match &v {
            serde_json::Value::Array(a) => a.iter().map(|x| scalar(&k, x, true)).collect::<Result<Vec<_>, _>>()?,
            other => vec![scalar(&k, other, false)?],
        };
        inputs.insert(k, vals);
    }
    Ok(inputs)
}

/// A graph parsed, planned and uploaded once (the reference re-parses it on every call, lib.rs:129-130).
pub struct DeviceGraph {
    handle: *mut ffi::gw_graph_t,
    pub info: ffi::gw_graph_info_t,
    pub inputs: InputSignalsInfo,
}

unsafe impl Send for DeviceGraph {}
unsafe impl Sync for DeviceGraph {}

impl DeviceGraph {
    /// storage::deserialize_witnesscalc_graph (storage.rs:214-249) + plan compilation
    pub fn load(graph_data: &[u8]) -> Result<DeviceGraph, Error> {
        let mut h: *mut ffi::gw_graph_t = std::ptr::null_mut();
        let mut st = ffi::gw_status_t { code: 0, error_msg: std::ptr::null_mut() };
        if unsafe { ffi::gw_graph_load(graph_data.as_ptr() as *const c_void, graph_data.len(), &mut h, &mut st) } != 0 {
            return Err(Error::Device(take_status(&mut st)));
        }
        let mut info = ffi::gw_graph_info_t::default();
        unsafe { ffi::gw_graph_info(h, &mut info) };
        let mut inputs = InputSignalsInfo::new();
        for i in 0..info.n_input_signals {
            let (mut name, mut off, mut len) = (std::ptr::null(), 0u32, 0u32);
            unsafe { ffi::gw_graph_input_signal(h, i, &mut name, &mut off, &mut len) };
            inputs.insert(unsafe { CStr::from_ptr(name) }.to_string_lossy().into_owned(), (off as usize, len as usize));
        }
        Ok(DeviceGraph { handle: h, info, inputs })
    }

    /// get_inputs_buffer + populate_inputs (lib.rs:154-181) for one input set: slot 0 = 1, missing keys stay 0
    pub fn inputs_buffer(&self, inputs: &HashMap<String, Vec<U256>>) -> Result<Vec<U256>, Error> {
        let mut buf = vec![U256::ZERO; self.info.n_inputs as usize];
        buf[0] = U256::from(1u64);
        for (k, vals) in inputs {
            let (off, len) = *self.inputs.get(k).ok_or_else(|| Error::Device(format!("unknown input signal {}", k)))?;
            if len != vals.len() {
                return Err(Error::Device(format!("Invalid input length for {}", k)));
            }
            buf[off..off + len].copy_from_slice(vals);
        }
        Ok(buf)
    }

    /// graph::evaluate (graph.rs:367-391) for many input sets on `n_gpus` devices: one witness per set, identical to
    /// what `evaluate(nodes, &inputs[k], outputs)` returns.
    pub fn evaluate_batch(&self, inputs: &[Vec<U256>], n_gpus: i32) -> Result<Vec<Vec<U256>>, Error> {
        let (i, w) = (self.info.n_inputs as usize, self.info.n_witness as usize);
        let mut inb = vec![0u8; inputs.len() * i * 32];
        for (k, row) in inputs.iter().enumerate() {
            for (j, v) in row.iter().enumerate() {
                inb[(k * i + j) * 32..(k * i + j) * 32 + 32].copy_from_slice(&v.to_le_bytes::<32>());
            }
        }
        let mut out = vec![0u8; inputs.len() * w * 32];
        let mut st = ffi::gw_status_t { code: 0, error_msg: std::ptr::null_mut() };
        let rc = unsafe {
            ffi::gw_calc_witness_batch(self.handle, inb.as_ptr(), inputs.len(), out.as_mut_ptr(), std::ptr::null_mut(), n_gpus, &mut st)
        };
        if rc != 0 {
            return Err(Error::Device(take_status(&mut st)));
        }
        Ok(out.chunks_exact(w * 32).map(|r| r.chunks_exact(32).map(U256::from_le_slice).collect()).collect())
    }

    /// Streaming variant: `consumer(first_set, rows)` sees chunks of packed witness rows (n x W x 32 bytes) as they land
    /// in the library's pinned ring; nothing the size of the whole batch is ever allocated.
    pub fn evaluate_stream<F>(&self, packed_inputs: &[u8], n_sets: usize, n_gpus: i32, mut consumer: F) -> Result<(), Error>
    where
        F: FnMut(usize, &[u8]) -> bool + Send,
    {
        unsafe extern "C" fn tramp<F: FnMut(usize, &[u8]) -> bool>(
            user: *mut c_void, _device: i32, first: usize, n: usize, rows: *const u8, row_bytes: usize, _flags: *const u32,
        ) -> i32 {
            // one worker thread per GPU may call this concurrently: serialise the FnMut
            static LOCK: std::sync::Mutex<()> = std::sync::Mutex::new(());
            let _g = LOCK.lock().unwrap();
            let f = &mut *(user as *mut F);
            if f(first, std::slice::from_raw_parts(rows, n * row_bytes)) { 0 } else { 1 }
        }
        let mut st = ffi::gw_status_t { code: 0, error_msg: std::ptr::null_mut() };
        let rc = unsafe {
            ffi::gw_calc_witness_batch_stream(self.handle, 0, n_gpus, packed_inputs.as_ptr(), n_sets, 0, tramp::<F>,
                                              &mut consumer as *mut F as *mut c_void, &mut st)
        };
        if rc != 0 {
            return Err(Error::Device(take_status(&mut st)));
        }
        Ok(())
    }
}

impl Drop for DeviceGraph {
    fn drop(&mut self) {
        unsafe { ffi::gw_graph_free(self.handle) }
    }
}
