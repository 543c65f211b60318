//! `graph::evaluate` (src/graph.rs:367-391) over the device library.
//!
//! The reference's `evaluate(nodes: &[Node], inputs: &[U256], outputs: &[usize]) -> Vec<U256>` walks an in-memory
//! node list.  The device library consumes the SERIALISED graph (the `wtns.graph.001` bytes that
//! `storage::serialize_witnesscalc_graph` writes, storage.rs:137-183), so the shim's `evaluate` takes a `DeviceGraph`
//! (load once) instead of `&[Node]`; a caller that only has `Vec<Node>` serialises it with the reference's own
//! `storage::serialize_witnesscalc_graph` first.
use ruint::aliases::U256;

use crate::{DeviceGraph, Error};

/// one input set: `inputs` is the flat buffer of `get_inputs_buffer`/`populate_inputs` (slot 0 is forced to 1)
pub fn evaluate(graph: &DeviceGraph, inputs: &[U256]) -> Result<Vec<U256>, Error> {
    Ok(graph.evaluate_batch(&[inputs.to_vec()], 1)?.remove(0))
}

/// many input sets, sharded over `n_gpus` devices (no collective: witnesses are independent)
pub fn evaluate_batch(graph: &DeviceGraph, inputs: &[Vec<U256>], n_gpus: i32) -> Result<Vec<Vec<U256>>, Error> {
    graph.evaluate_batch(inputs, n_gpus)
}
