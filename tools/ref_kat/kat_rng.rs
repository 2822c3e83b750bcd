//! Scripted stand-in for `rand::rngs::ThreadRng`, added to the COPY of the reference crate by tools/ref_kat/run.sh so that
//! the functions that draw random numbers (`disney_sample`, `sample_light`, `direct_light`; tracer.rs:137,191-192,446-447,534)
//! can be called on chosen draws.  `rng.gen::<f32>()` is rand 0.8.5's `Standard` distribution: `(next_u32() >> 8) as f32 *
//! 2^-24` — so a draw u on the 2^-24 grid is scripted as the word `(u * 2^24) << 8`.  Nothing in tracer.rs changes except the
//! `use` line that names the type.
use rand::RngCore;
use std::collections::VecDeque;
use std::sync::Mutex;

// process-wide, not thread-local: Tracer::render runs its pixel loop on a rayon worker thread (tracer.rs:29-32)
static SCRIPT: Mutex<VecDeque<u32>> = Mutex::new(VecDeque::new());

/// Queue the next draws (each in [0, 1) on the 2^-24 grid); replaces what is left of the previous script.
pub fn script(draws: &[f32]) {
    let mut s = SCRIPT.lock().unwrap();
    s.clear();
    for u in draws {
        let w = (*u as f64 * 16777216.0) as u32;
        s.push_back(w << 8);
    }
}
/// Draws left unconsumed (a KAT records how many draws a call consumed).
pub fn remaining() -> usize { SCRIPT.lock().unwrap().len() }

pub struct ThreadRng;
pub fn thread_rng() -> ThreadRng { ThreadRng }

impl RngCore for ThreadRng {
    fn next_u32(&mut self) -> u32 { SCRIPT.lock().unwrap().pop_front().expect("kat_rng: script exhausted") }
    fn next_u64(&mut self) -> u64 { let lo = self.next_u32() as u64; let hi = self.next_u32() as u64; (hi << 32) | lo }
    fn fill_bytes(&mut self, dest: &mut [u8]) { for b in dest.iter_mut() { *b = (self.next_u32() >> 24) as u8; } }
    fn try_fill_bytes(&mut self, dest: &mut [u8]) -> Result<(), rand::Error> { self.fill_bytes(dest); Ok(()) }
}
