"""Linear-algebra helpers of the Fisher path (reference: curvature/utils.py:21-60, 288-310)."""
from typing import Dict, List

import torch
from torch import Tensor
from torch.nn import Module


def get_eigenvectors(factors: Dict[Module, Tensor]) -> Dict[Module, Tensor]:
    """Eigenvectors (columns, ascending eigenvalues) of `F + F^T` for both Kronecker factors of every layer
    (reference: utils.py:45-60, which calls the removed `torch.symeig`).  One-shot: this is the single place
    where a library eigensolver (cuSOLVER syevd through `torch.linalg.eigh`) is used; it is timed separately
    from the estimation pass."""
    eigenvectors = dict()
    for layer, (xxt, ggt) in factors.items():
        sym_xxt, sym_ggt = xxt + xxt.t(), ggt + ggt.t()
        _, xxt_eigvecs = torch.linalg.eigh(sym_xxt, UPLO='U')
        _, ggt_eigvecs = torch.linalg.eigh(sym_ggt, UPLO='U')
        eigenvectors[layer] = (xxt_eigvecs.contiguous(), ggt_eigvecs.contiguous())
    return eigenvectors


def eigendecompose(factors: Dict[Module, Tensor]):
    """ONE eigensolve per factor for both consumers of the one-shot eigenbases: returns
    `(eigvecs, eigvals)` with `eigvecs[layer] = (QA, QG)` exactly as `get_eigenvectors` (eigenvectors of `F + F^T`,
    utils.py:45-60) and `eigvals[layer] = (wA, wG)` the spectra of `F` itself (what `get_eigenvalues`, utils.py:21-42,
    computes with a second `symeig`): for the symmetric factors this path produces, `eig(F) = eig(F + F^T) / 2`."""
    eigvecs, eigvals = dict(), dict()
    for layer, (xxt, ggt) in factors.items():
        wa, qa = torch.linalg.eigh(xxt + xxt.t(), UPLO='U')
        wg, qg = torch.linalg.eigh(ggt + ggt.t(), UPLO='U')
        eigvecs[layer] = (qa.contiguous(), qg.contiguous())
        eigvals[layer] = (wa / 2, wg / 2)
    return eigvecs, eigvals


def get_eigenvalues(factors: List[Tensor],
                    verbose: bool = False,
                    eigvals: List = None) -> Tensor:
    """Eigenvalues of KFAC, EFB or diagonal factors (reference: utils.py:21-42): for a pair of Kronecker factors
    the outer product of their spectra, otherwise the factor flattened.  `eigvals` optionally supplies the spectra
    already computed by `eigendecompose` (one per entry of `factors`, in order): no second eigensolve."""
    chunks = []
    for layer, factor in enumerate(factors):
        if verbose:
            print(f"Layer [{layer + 1}/{len(factors)}]")
        if len(factor) == 2 and eigvals is not None:
            chunks.append(torch.outer(eigvals[layer][0], eigvals[layer][1]).contiguous().view(-1))
        elif len(factor) == 2:
            xxt_eigvals = torch.linalg.eigvalsh(factor[0], UPLO='U')
            ggt_eigvals = torch.linalg.eigvalsh(factor[1], UPLO='U')
            chunks.append(torch.outer(xxt_eigvals, ggt_eigvals).contiguous().view(-1))
        else:
            chunks.append(factor.contiguous().view(-1))
    if not chunks:
        return Tensor()
    return torch.cat([c.to(chunks[0].device) for c in chunks])


def kron(a: Tensor,
         b: Tensor) -> Tensor:
    r"""Kronecker product of two 2-D tensors (reference: utils.py:288-310).

    Examples:
        >>> a = torch.tensor([[1, 2], [3, 4]])
        >>> b = torch.tensor([[0, 5], [6, 7]])
        >>> kron(a, b)
        tensor([[ 0,  5,  0, 10],
                [ 6,  7, 12, 14],
                [ 0, 15,  0, 20],
                [18, 21, 24, 28]])
    """
    return (a[:, None, :, None] * b[None, :, None, :]).reshape(a.size(0) * b.size(0), a.size(1) * b.size(1))
